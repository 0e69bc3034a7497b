"""EXPERIMENTAL: run the tensor-core forward (csrc/forward_tc.cu, error-compensated TF32 on tcgen05) next to the product
kernel on the same packed model and pre-processed inputs: parity under the oracle's rule, and time per forward.
    python tools/tc_forward.py [B=4096]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from monoloco_b200 import synthetic, _lib as L_
from monoloco_b200.engine import LocoEngine
from oracle import loco_oracle as O


class TcEngine:
    def __init__(self, eng, max_rows):
        self.lib, pm = L_.lib(), eng.packed
        d = pm.desc
        desc = L_.MlbModelDesc(L_.MLB_ABI_VERSION, d['input_size'], d['output_size'], d['linear_size'], d['n_ops'], d['decode_kind'],
                               d['p_dropout'], 0)
        ops = (L_.MlbOp * len(pm.ops))()
        for i, o in enumerate(pm.ops):
            ops[i] = L_.MlbOp(o['type'], o['K'], o['Kpad'], o['N'], o['flags'], o['out_col'], o['w_off'], o['scale_off'], o['shift_off'])
        self.h = C.c_void_p()
        self.lib.mlb_tc_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        self.lib.mlb_tc_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.mlb_tc_destroy.argtypes = [C.c_void_p]
        L_.check(self.lib.mlb_tc_create(C.byref(desc), ops, pm.blob.ctypes.data_as(C.c_void_p), pm.blob.size, eng.index, max_rows,
                                        C.byref(self.h)), 'mlb_tc_create')
        self.out_size = d['output_size']

    def forward(self, x):
        out = torch.zeros((x.shape[0], self.out_size), dtype=torch.float32, device=x.device)
        L_.check(self.lib.mlb_tc_forward(self.h, x.data_ptr(), x.shape[0], out.data_ptr(), None), 'mlb_tc_forward')
        return out

    def close(self):
        self.lib.mlb_tc_destroy(self.h)


def compare(B=4096, kind='loco', reps=20):
    sd = synthetic.make_state_dict(kind, 34, 9, 1024, 3, 0)
    eng = LocoEngine(sd)
    tc = TcEngine(eng, B)
    x = torch.from_numpy(synthetic.make_inputs(B, 34, seed=1)).cuda()
    ref = eng.forward(x)['raw']
    got = tc.forward(x)
    torch.cuda.synchronize()
    ok, worst = O.close(got.cpu().numpy(), ref.cpu().numpy())
    idx = np.random.RandomState(0).choice(B, min(B, 256), replace=False)
    ok_o, worst_o = O.close(got.cpu().numpy()[idx], O.model_forward(sd, x.cpu().numpy()[idx]))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_tc, t_ff = timed(lambda: tc.forward(x)), timed(lambda: eng.forward(x, kernel='tile'))
    print('B=%d %s: tensor-core path %.3f ms, FFMA row-tile kernel %.3f ms (x%.2f); vs FFMA kernel close=%s worst/tol=%.3f; '
          'vs oracle close=%s worst/tol=%.3f' % (B, kind, t_tc, t_ff, t_ff / t_tc, ok, worst, ok_o, worst_o))
    tc.close()
    eng.close()
    return ok_o, worst_o, t_tc, t_ff


if __name__ == '__main__':
    compare(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
