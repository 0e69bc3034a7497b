"""Checkpoint -> device blob + layer program.

Turns a reference state_dict (key names: SURVEY.md §8b "Checkpoint ABI";
monoloco/network/architectures.py:8-46, 111-133) into
  * one contiguous fp32 blob: per L-wide Linear the transposed weight W^T [Kpad, L] (so that a chunk
    of KC consecutive k rows is one contiguous TMA bulk copy), the eval-mode BatchNorm1d folded into a
    per-feature (scale, shift) pair, and the narrow head weights in their native [N, K] layout;
  * the op table of include/monoloco_b200.h (`mlb_op`).
BatchNorm folding: y = ((Wx + b) - mean) / sqrt(var + eps) * gamma + beta = (Wx) * s + ((b - mean) * s + beta),
s = gamma / sqrt(var + eps), computed in float64 and rounded once to fp32.
"""
import numpy as np

from . import _lib as L_

KC = 8          # must match csrc/common.cuh
BN_EPS = 1e-5   # nn.BatchNorm1d default (architectures.py:25)
ALIGN = 32      # floats (128 B): TMA bulk copies need 16 B, keep cache-line alignment


def _np(v):
    if hasattr(v, 'detach'):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def num_stages(sd):
    n = 0
    while 'linear_stages.%d.w1.weight' % n in sd:
        n += 1
    return n


def padded_width(linear_size):
    """Hidden width the kernels run: the next multiple of 128 up to 1024 (FFMA kernels, and the tensor-core kernel when it
    is also a multiple of 256), the next multiple of 256 up to 2048 beyond (tensor-core kernel only).  `--hidden_size`
    is free in the reference (run.py:101,122; hyp_tuning.py:52 uses 2048)."""
    if linear_size <= 1024:
        return ((linear_size + 127) // 128) * 128
    if linear_size <= 2048:
        return ((linear_size + 255) // 256) * 256
    raise ValueError("monoloco_b200: linear_size up to 2048 is supported (got %d)" % linear_size)


class PackedModel:
    def __init__(self, desc, ops, blob, kind):
        self.desc, self.ops, self.blob, self.kind = desc, ops, blob, kind


def pack_state_dict(sd, p_dropout=0.2):
    sd = {k: _np(v) for k, v in sd.items()}
    is_loco = 'w_fin.weight' in sd
    L_real, in_size = sd['w1.weight'].shape
    L = padded_width(L_real)   # hidden units beyond L_real are zero columns / rows: they stay exactly 0 through every layer
    n_stage = num_stages(sd)
    chunks = []
    cursor = [0]

    def put(arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        off = cursor[0]
        chunks.append((off, arr))
        cursor[0] = off + ((arr.size + ALIGN - 1) // ALIGN) * ALIGN
        return off

    def affine(lin, bn):
        b = sd[lin + '.bias'].astype(np.float64)
        if bn is None:
            s, t = np.ones_like(b), b
        else:
            s = sd[bn + '.weight'].astype(np.float64) / np.sqrt(sd[bn + '.running_var'].astype(np.float64) + BN_EPS)
            t = (b - sd[bn + '.running_mean'].astype(np.float64)) * s + sd[bn + '.bias'].astype(np.float64)
        pad = L - s.size   # padded units: scale 1, shift 0 on an all-zero weight column -> 0 (and ReLU(0) = 0)
        return np.concatenate([s, np.ones(pad)]), np.concatenate([t, np.zeros(pad)])

    ops = []

    def gemm(lin, bn, flags):
        w = sd[lin + '.weight']  # [N, K]
        n, k = w.shape
        assert n == L_real
        k_op = L if k == L_real and not (flags & L_.F_IN_XIN) else k   # hidden layers: K = padded width
        kpad = ((k_op + KC - 1) // KC) * KC
        wt = np.zeros((kpad, L), dtype=np.float32)
        wt[:k, :n] = w.T
        s, t = affine(lin, bn)
        ops.append(dict(type=L_.OP_GEMM, K=k_op, Kpad=kpad, N=L, flags=flags, out_col=0,
                        w_off=put(wt), scale_off=put(s), shift_off=put(t)))

    def head(lin, out_col):
        w = sd[lin + '.weight']
        n, k = w.shape
        assert k == L_real
        wp = np.zeros((n, L), dtype=np.float32)
        wp[:, :k] = w
        ops.append(dict(type=L_.OP_HEAD, K=L, Kpad=L, N=n, flags=0, out_col=out_col,
                        w_off=put(wp), scale_off=0, shift_off=put(sd[lin + '.bias'])))

    gemm('w1', 'batch_norm1', L_.F_RELU | L_.F_DROPOUT | L_.F_IN_XIN | (L_.F_SAVE_RES if n_stage else 0))
    for i in range(n_stage):
        p = 'linear_stages.%d' % i
        gemm(p + '.w1', p + '.batch_norm1', L_.F_RELU)
        gemm(p + '.w2', p + '.batch_norm2', L_.F_RELU | L_.F_ADD_RES | (L_.F_SAVE_RES if i + 1 < n_stage else 0))
    if is_loco:
        out_size = sd['w_fin.weight'].shape[0] + 1  # architectures.py:14,42,70
        gemm('w2', None, 0)
        head('w_aux', out_size - 1)
        gemm('w3', 'batch_norm3', L_.F_RELU | L_.F_DROPOUT)
        head('w_fin', 0)
        decode = L_.DECODE_LOCO if out_size in (9, 10) else L_.DECODE_NONE
    else:
        out_size = sd['w2.weight'].shape[0]
        head('w2', 0)
        decode = {9: L_.DECODE_MONO, 2: L_.DECODE_DB}.get(out_size, L_.DECODE_NONE)

    blob = np.zeros(cursor[0], dtype=np.float32)
    for off, arr in chunks:
        blob[off:off + arr.size] = arr
    desc = dict(input_size=int(in_size), output_size=int(out_size), linear_size=int(L), n_ops=len(ops),
                decode_kind=decode, p_dropout=float(p_dropout))
    return PackedModel(desc, ops, blob, 'loco' if is_loco else 'monoloco')


def flops_per_detection(sd):
    """2 * sum(in*out) over every Linear (SURVEY.md §8d 'Algorithmic flops')."""
    return int(sum(2 * _np(v).size for k, v in sd.items() if k.endswith('.weight') and _np(v).ndim == 2))


def weight_bytes(sd):
    """fp32 parameter bytes (Linear weights+biases, BN affine) + BN running stats."""
    return int(sum(4 * _np(v).size for k, v in sd.items() if not k.endswith('num_batches_tracked')))
