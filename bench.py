#!/usr/bin/env python
"""
bench.py -- monoloco hot path on B200:  detections/s of the fused forward (pre-process -> LocoModel -> decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Contract (see the task statement / DESIGN.md §6):
  * one step = one pass of the hot path over one batch of synthetic 17-keypoint detections
    (BASELINE.json configs[1]: LocoModel mono 34->9, 3 stages x 1024, batch 4096 per GPU, fp32);
  * `value`  = whole-job detections/s with inputs resident in HBM, timed with CUDA events on the launching stream,
               L2 flushed (256 MiB memset) before every timed step, max over ranks;
  * `e2e`    = same metric through the C-ABI host-buffer call (pinned host memory; H2D + kernel + D2H per step);
  * `roofline`, `cpu_baseline`, `clocks`, `gpu_launches` as specified.
  * --impl reference: the reference's CPU implementation of the path (oracle/torch_port.py = the same torch-eager
    op sequence as the reference nn.Module) timed on the host cores.
Multi-GPU: launched by torch.distributed.run; detections shard over ranks (weak scaling, 4096 per GPU), one
all-gather of the [B,17] outputs per step (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "detections/sec LocoModel(34->9, 3x1024) fused forward @ batch 4096 per GPU"
UNIT = "detections/s"


def measured_traffic():
    """dram__bytes_read+write of the forward kernel from the committed ncu capture (profiles/forward_traffic.json)."""
    path = os.path.join(ROOT, 'profiles', 'forward_traffic.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get('traffic_bytes'), d.get('source')
    return None, None


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), 'measured', d
    return 6650.0, 'fallback', {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md): one persistent
    `nvidia-smi -lms 20` process, started before and killed after the region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # first sample lands before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        self.rows = [[c.strip() for c in ln.split(',')] for ln in out.strip().splitlines() if ln.strip()]

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def best_cpu_threads(fn, budget_s=6.0):
    """The reference's CPU path is torch eager; on a many-core shared host `all threads` is often NOT its fastest
    setting (128 threads on the GPU box: 2.4 s per 4096-batch vs 0.1-0.2 s at 16-32).  Give the baseline its best
    thread count: try a few, keep the fastest, and report that count as `cores`."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], None
    t_all = time.perf_counter()
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
        if time.perf_counter() - t_all > budget_s:
            break
    torch.set_num_threads(best)
    return best


def cpu_reference_rate(sd, x_np, budget_s=12.0, min_reps=3):
    """detections/s of the reference's CPU path (torch eager, best host thread count) on a bounded sample."""
    from oracle import torch_port as T  # the one place bench.py executes oracle/: the timed CPU baseline
    tsd = T.to_torch(sd)
    x = torch.from_numpy(x_np)
    with torch.no_grad():
        best_cpu_threads(lambda: T.model_forward(tsd, x))
        for _ in range(2):
            T.model_forward(tsd, x)
        times = []
        t_all = time.perf_counter()
        while len(times) < min_reps or (time.perf_counter() - t_all < budget_s and len(times) < 200):
            t0 = time.perf_counter()
            T.model_forward(tsd, x)
            times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return x_np.shape[0] / med, len(times), med


def extras(eng, sd, dev):
    """Side measurements outside the timed region (other BASELINE.json configs), device-timed, L2 warm:
    small-batch latency (configs[0]/[1]: batch 1 / 256 -> cluster kernel), batch 65536, stereo 64x64 pairs + filter
    (configs[2]), and the one-launch training step at batch 4096 (configs[3])."""
    from monoloco_b200 import synthetic, _lib as L_
    out = {}

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    try:
        lat = {}
        for b in (1, 16, 256, 65536):
            x = torch.from_numpy(synthetic.make_keypoints(b, seed=2)).to(dev)
            lat[str(b)] = timed(lambda: eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS), 20 if b <= 4096 else 3)
        out["forward_ms_by_batch"] = lat
        # one image's worth of detections is the weight-streaming regime (B <= ~19): the whole-grid kernel's
        # share of the HBM copy peak on the 33.8 MB it has to touch (L2-warm here, so this is an L2/HBM mix)
        hbm_peak = peaks()[0]
        wbytes = eng.packed.blob.size * 4
        out["small_batch_roofline"] = {"rows": 16, "kernel": "loco_forward_wide_kernel", "ms": lat["16"],
                                       "weight_bytes": wbytes, "achieved_GBps": wbytes / (lat["16"] * 1e-3) / 1e9,
                                       "frac_of_hbm_copy_peak": wbytes / (lat["16"] * 1e-3) / 1e9 / hbm_peak}
        from monoloco_b200 import engine as E
        seng = E.LocoEngine(synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2), device=dev)
        le, ri = synthetic.make_keypoints(64, seed=3, right=True)
        le, ri = torch.from_numpy(le).to(dev), torch.from_numpy(ri).to(dev)

        def stereo():
            o = seng.forward(le, x_right=ri, kk=synthetic.KITTI_K, kind=L_.IN_KPS_STEREO, want_xyzc=True)
            seng.stereo_filter(o['raw'], o['dec'], 64, 64, xyzc=o['xyzc'])
        out["stereo_64x64_pairs_plus_filter_ms"] = timed(stereo, 10)
        seng.close()
        from monoloco_b200.network.architectures import LocoModel
        from monoloco_b200.train import train_step
        m = LocoModel(34, 9, 1024, p_dropout=0.2, num_stage=3)
        m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        m.to(dev).train()
        xt = torch.from_numpy(synthetic.make_inputs(4096, 34, seed=3)).to(dev)
        yt = torch.from_numpy(synthetic.make_labels(4096, seed=4)).to(dev)
        tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori')
        ms = timed(lambda: train_step(m, xt, yt, tasks), 5)
        out["train_step_b4096"] = {"ms": ms, "tflops": 3 * 16865280 * 4096 / ms / 1e9,
                                   "what": "forward + MultiTaskLoss + backward + dW, one cooperative launch, fp32"}
    except Exception as exc:  # side measurements must never break the contract line
        out["error"] = repr(exc)
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores."""
    if rank != 0:
        return
    from monoloco_b200 import synthetic
    from oracle import loco_oracle as O
    from oracle import torch_port as T
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    B = args.batch
    kps = synthetic.make_keypoints(B, seed=0)
    tsd = T.to_torch(sd)
    times = []
    with torch.no_grad():
        x0 = torch.from_numpy(O.preprocess_monoloco(kps, synthetic.KITTI_K))
        best_cpu_threads(lambda: T.model_forward(tsd, x0))
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            x = O.preprocess_monoloco(kps, synthetic.KITTI_K)           # process.py:47-67
            out = T.model_forward(tsd, torch.from_numpy(x))            # architectures.py:48-71
            O.extract_outputs(out.numpy())                             # process.py:231-278
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    val = B / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LocoModel mono 34->9 L=1024 x3 stages, pre-process + forward + decode, batch %d, CPU" % B,
                       "batch_per_step": B},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "host_cpus": os.cpu_count(),
                             "sample": "%d steps x %d detections, torch-eager CPU restatement (oracle/torch_port.py), "
                                       "best of several torch thread counts" % (args.steps, B)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=4096, help='detections per GPU per step')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the small-batch latency / train-step side measurements')
    ap.add_argument('--rows-per-group', type=int, default=0)
    ap.add_argument('--gather', default='fused', choices=['fused', 'nccl'],
                    help='multi-GPU output all-gather: fused peer stores from the kernel epilogue, or NCCL')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from monoloco_b200 import synthetic, engine, packing, _lib as L_

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B = args.batch
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd, device=dev)
    lib = L_.lib()
    kk = synthetic.KITTI_K
    kps_host = torch.from_numpy(synthetic.make_keypoints(B, seed=rank)).pin_memory()
    kps = kps_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    st = torch.cuda.current_stream(dev)
    sharded = None
    if world > 1:
        from monoloco_b200 import distributed as D
        sharded = D.ShardedLoco(eng, world * B, mode=args.gather)

    def step():
        if sharded is not None:
            return sharded.forward(kps, kk, rows_per_group=args.rows_per_group)  # forward + all-gather of [N*B, 20] rows
        return eng.forward(kps, kk=kk, kind=L_.IN_KPS, rows_per_group=args.rows_per_group)

    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- timed region: K steps, device-timed, L2 flushed before each
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.mlb_launch_count()
    with ClockSampler(local_rank) as clocks:
        for e0, e1 in ev:
            flush.zero_()
            e0.record(st)
            step()
            e1.record(st)
        torch.cuda.synchronize(dev)
    launches = lib.mlb_launch_count() - launches0
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
    total_ms = float(sum(e0.elapsed_time(e1) for e0, e1 in ev))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---------------- e2e through the host-buffer C-ABI call (pinned host memory)
    out_host = {'raw': torch.empty((B, 9)).pin_memory(), 'dec': torch.empty((B, 8)).pin_memory()}
    for _ in range(3):
        eng.forward_host(kps_host, kk=kk, kind=L_.IN_KPS, out=out_host)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.forward_host(kps_host, kk=kk, kind=L_.IN_KPS, out=out_host)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms * 1e-3)

    if rank == 0:
        hbm_peak, peak_src, pk = peaks()
        w_bytes = eng.packed.blob.size * 4
        io_bytes = 204 + 36 + 32          # raw kps in + raw out + decoded out
        alg_bytes = w_bytes + B * io_bytes
        flops = packing.flops_per_detection(sd) * B
        # single-GPU kernel time: at world == 1 the step is exactly one launch of loco_forward_kernel
        achieved_gbs = alg_bytes / (ms_per_step * 1e-3) / 1e9
        ffma_peak = engine.probe_ffma_tflops(local_rank)
        achieved_tf = flops / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LocoModel mono 34->9 L=1024 x3 stages: raw keypoints [B,3,17] -> fused pre-process + "
                                   "forward + decode, batch %d per GPU" % B,
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": "dp%d" % world,
                       "l2": "flushed before every timed step (256 MiB memset)",
                       "collective": ("none" if world == 1 else
                                      ("kernel-epilogue peer stores over NVLink (cudaIpc) + barrier" if args.gather == 'fused'
                                       else "NCCL all_gather_into_tensor") + " of [N*B,20] fp32 rows per step")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 51 * 4, "d2h_bytes_per_step": B * 17 * 4,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": measured_traffic()[0] if B == 4096 else None,
                         "traffic_source": measured_traffic()[1], "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes,
                         "note": "at batch 4096 the path is FP32-FFMA bound (SURVEY.md §0.4); see fp32",
                         "fp32": {"achieved": achieved_tf, "peak": ffma_peak, "unit": "TFLOP/s",
                                  "frac": achieved_tf / ffma_peak if ffma_peak else None,
                                  "peak_source": "measured in-run by mlb_probe_ffma (pure FFMA kernel)",
                                  "algorithmic_flops": flops}},
        }
        if not args.no_extras and world == 1:
            line["extras"] = extras(eng, sd, dev)
        if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only
            x = np.ascontiguousarray(synthetic.make_inputs(B, 34, seed=0))
            rate, reps, med = cpu_reference_rate(sd, x)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "host_cpus": os.cpu_count(),
                                    "sample": "%d x batch-%d model forwards (oracle/torch_port.py), median %.1f ms, best of "
                                              "several torch thread counts" % (reps, B, med * 1e3)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
