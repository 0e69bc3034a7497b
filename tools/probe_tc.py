"""Run the tcgen05 TF32x3 probe (csrc/probe_tc.cu) on cuda:0 and compare it with the CPU emulations of
tools/tf32x3_study.py: which accumulator rounding model does the tensor core follow, and how far is each mode from fp64?
    python tools/probe_tc.py [K=1024]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from monoloco_b200 import _lib as L_
from tools.tf32x3_study import to_tf32, mm_tf32x3, mm_tf32x3_split


def run(A, W, mode):
    lib = L_.lib()
    lib.mlb_probe_tf32x3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    a, w = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    main = torch.zeros((128, 128), dtype=torch.float32, device='cuda')
    cross = torch.zeros((128, 128), dtype=torch.float32, device='cuda')
    L_.check(lib.mlb_probe_tf32x3(a.data_ptr(), w.data_ptr(), A.shape[1], mode, main.data_ptr(), cross.data_ptr(), None), 'probe')
    torch.cuda.synchronize()
    return main.cpu().numpy(), cross.cpu().numpy()


def run_layer(B=4096, N=1024, K=1024, reps=20):
    """One whole layer through mlb_probe_tc_layer: accuracy vs fp64 on a row sample, GEMM time by CUDA events."""
    lib = L_.lib()
    lib.mlb_probe_tc_layer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p]
    rng = np.random.RandomState(2)
    X = np.abs(rng.standard_normal((B, K))).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    x, w = torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda()
    y = torch.zeros((B, N), dtype=torch.float32, device='cuda')
    xp = torch.empty(2 * B * K, dtype=torch.float32, device='cuda')
    wp = torch.empty(2 * N * K, dtype=torch.float32, device='cuda')
    call = lambda stages: L_.check(lib.mlb_probe_tc_layer(x.data_ptr(), w.data_ptr(), y.data_ptr(), B, N, K, xp.data_ptr(),  # noqa: E731
                                                          wp.data_ptr(), stages, None), 'layer')
    call(7)
    torch.cuda.synchronize()
    rows = rng.choice(B, 64, replace=False)
    ref = X[rows].astype(np.float64) @ W.astype(np.float64).T
    got = y[torch.from_numpy(rows).cuda()].cpu().numpy()
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    sg = float(np.abs((X[rows] @ W.T) - ref).max() / np.abs(ref).max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call(4)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print('layer %dx%dx%d on tcgen05 (3xTF32, two accumulators): %.1f us = %.1f TFLOP/s fp32-equivalent; max err %.2e (numpy sgemm %.2e)'
          % (B, N, K, us, 2.0 * B * N * K / us / 1e6, err, sg))
    return us, err


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'layer':
        run_layer(*[int(v) for v in sys.argv[2:]])
        sys.exit(0)
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    rng = np.random.RandomState(0)
    A = np.abs(rng.standard_normal((128, K))).astype(np.float32)          # post-ReLU-like activations (same sign: worst case
    W = (rng.standard_normal((128, K)) / np.sqrt(K)).astype(np.float32)   # for truncation bias), weights ~ N(0, 1/K)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(ref).max()
    rel = lambda x: float(np.abs(x - ref).max() / scale)  # noqa: E731
    print('fp32 (numpy sgemm)          max|err|/max|ref| = %.3e' % rel((A @ W.T).astype(np.float64)))
    m0, _ = run(A, W, 0)
    print('tcgen05 1xTF32              %.3e   (emulation %.3e)' % (rel(m0), rel(to_tf32(A).astype(np.float64) @ to_tf32(W).astype(np.float64).T)))
    m1, _ = run(A, W, 1)
    print('tcgen05 3xTF32, one acc     %.3e   (emulation rn %.3e, rz %.3e)' % (rel(m1), rel(mm_tf32x3(A, W, 'rn')), rel(mm_tf32x3(A, W, 'rz'))))
    m2, c2 = run(A, W, 2)
    print('tcgen05 3xTF32, cross apart %.3e   (emulation rn %.3e, rz %.3e)' % (rel(m2 + c2), rel(mm_tf32x3_split(A, W, 'rn')), rel(mm_tf32x3_split(A, W, 'rz'))))
    for name, got, emu in (('one acc', m1, mm_tf32x3), ('cross apart', m2 + c2, mm_tf32x3_split)):
        d_rn = float(np.abs(got - emu(A, W, 'rn')).max() / scale)
        d_rz = float(np.abs(got - emu(A, W, 'rz')).max() / scale)
        print('   %-12s distance to the rn emulation %.3e, to the rz emulation %.3e' % (name, d_rn, d_rz))
