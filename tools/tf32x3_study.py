"""Feasibility study for a tensor-core path (round 2 candidate): can error-compensated TF32 (a = a_hi + a_lo, 3 MMAs per
product, fp32 accumulation per k=8 MMA block) keep the LocoModel forward inside the 1e-5 parity rule?  Pure numpy
emulation on the CPU -- no GPU needed.  Accumulator rounding per MMA is emulated as round-to-nearest ('rn') and as
truncation ('rz', the pessimistic model of the tensor-core adder)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from monoloco_b200 import synthetic
from oracle import loco_oracle as O


def to_tf32(x):
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest (ties away from zero)."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return r.view(np.float32)


def rz32(x64):
    """float64 -> float32 with truncation toward zero."""
    y = x64.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y


def mm_tf32x3(a, w, mode):
    """a [B,K] @ w[N,K]^T with 3 TF32 MMAs per k-block of 8 and an fp32 accumulator rounded once per MMA."""
    ah, wh = to_tf32(a), to_tf32(w)
    al, wl = to_tf32(a - ah), to_tf32(w - wh)
    acc = np.zeros((a.shape[0], w.shape[0]), dtype=np.float32)
    K = a.shape[1]
    for k0 in range(0, K, 8):
        sl = slice(k0, min(K, k0 + 8))
        for x, y in ((al, wh), (ah, wl), (ah, wh)):  # small terms first
            p = x[:, sl].astype(np.float64) @ y[:, sl].astype(np.float64).T  # products exact, block sum ~exact
            s = acc.astype(np.float64) + p
            acc = s.astype(np.float32) if mode == 'rn' else rz32(s)
    return acc


def mm_tf32x3_split(a, w, mode, kparts=1):
    """Same, but the two small cross terms go to their own accumulator and K is cut into `kparts` independent
    accumulators that are added in fp32 (RN) at the end: fewer truncations of the large partial sums."""
    ah, wh = to_tf32(a), to_tf32(w)
    al, wl = to_tf32(a - ah), to_tf32(w - wh)
    K = a.shape[1]
    total = np.zeros((a.shape[0], w.shape[0]), dtype=np.float32)
    step = -(-K // kparts)
    step = -(-step // 8) * 8
    rnd = (lambda v: v.astype(np.float32)) if mode == 'rn' else rz32
    for p0 in range(0, K, step):
        big = np.zeros_like(total)
        small = np.zeros_like(total)
        for k0 in range(p0, min(K, p0 + step), 8):
            sl = slice(k0, min(K, k0 + 8))
            big = rnd(big.astype(np.float64) + ah[:, sl].astype(np.float64) @ wh[:, sl].astype(np.float64).T)
            small = rnd(small.astype(np.float64) + al[:, sl].astype(np.float64) @ wh[:, sl].astype(np.float64).T)
            small = rnd(small.astype(np.float64) + ah[:, sl].astype(np.float64) @ wl[:, sl].astype(np.float64).T)
        total = total + (big + small)
    return total


def forward(sd, x, mm):
    def lin(v, name):
        return mm(v, sd[name + '.weight']) + sd[name + '.bias']

    def bn(v, name):
        g, b = sd[name + '.weight'], sd[name + '.bias']
        m, var = sd[name + '.running_mean'], sd[name + '.running_var']
        return (v - m) / np.sqrt(var + np.float32(1e-5)) * g + b
    relu = lambda v: np.maximum(v, 0)  # noqa: E731
    y = relu(bn(lin(x, 'w1'), 'batch_norm1'))
    for i in range(O.num_stages(sd)):
        pre = 'linear_stages.%d.' % i
        z = relu(bn(lin(y, pre + 'w1'), pre + 'batch_norm1'))
        z = relu(bn(lin(z, pre + 'w2'), pre + 'batch_norm2'))
        y = y + z
    y = lin(y, 'w2')
    aux = lin(y, 'w_aux')
    y = relu(bn(lin(y, 'w3'), 'batch_norm3'))
    return np.concatenate([lin(y, 'w_fin'), aux], axis=1).astype(np.float32)


if __name__ == '__main__':
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    sd = {k: np.asarray(v) for k, v in synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0).items()}
    x = O.preprocess_monoloco(synthetic.make_keypoints(B, seed=0), synthetic.KITTI_K)
    ref32 = O.loco_model_forward(sd, x)
    sd64 = {k: v.astype(np.float64) if v.dtype == np.float32 else v for k, v in sd.items()}
    ref64 = forward(sd64, x.astype(np.float64), lambda a, w: a @ w.T)
    print('fp32 oracle vs fp64      : close=%s worst/tol=%.3f' % O.close(ref32, ref64.astype(np.float32)))
    for mode in ('rn', 'rz'):
        out = forward(sd, x, lambda a, w: mm_tf32x3(np.asarray(a, np.float32), w, mode))
        ok, worst = O.close(out, ref32)
        ok64, worst64 = O.close(out, ref64.astype(np.float32))
        print('3xTF32 accumulate %s     : vs fp32 oracle close=%s worst/tol=%.3f | vs fp64 close=%s worst/tol=%.3f'
              % (mode, ok, worst, ok64, worst64))
    for kparts in (1, 2, 4):
        out = forward(sd, x, lambda a, w: mm_tf32x3_split(np.asarray(a, np.float32), w, 'rz', kparts))
        print('3xTF32 rz, cross terms apart, %d K-parts: vs fp32 oracle close=%s worst/tol=%.3f' % ((kparts,) + O.close(out, ref32)))
    out1 = forward(sd, x, lambda a, w: (to_tf32(np.asarray(a, np.float32)).astype(np.float64) @ to_tf32(w).astype(np.float64).T).astype(np.float32))
    print('1xTF32 (plain)           : vs fp32 oracle close=%s worst/tol=%.3f' % O.close(out1, ref32))
