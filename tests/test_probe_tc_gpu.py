"""Tensor-core probes (csrc/probe_tc.cu): what the accumulator of tcgen05.mma kind::tf32 does to an error-compensated fp32
product -- the measurement the tensor-core forward kernel (csrc/forward_tc.cu) rests on.  Ran green on a B200 in round 2."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def test_tf32x3_probe_close_to_fp64():
    from tools.probe_tc import run
    rng = np.random.RandomState(1)
    K = 256
    A = rng.standard_normal((128, K)).astype(np.float32)
    W = (rng.standard_normal((128, K)) / np.sqrt(K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    scale = np.abs(ref).max()
    m0, _ = run(A, W, 0)
    assert np.abs(m0 - ref).max() / scale < 5e-3          # plain TF32: ~1e-3
    m2, c2 = run(A, W, 2)
    assert np.abs(m2 + c2 - ref).max() / scale < 5e-6     # error-compensated: fp32-class


def test_tc_layer_probe_close_to_fp64():
    from tools.probe_tc import run_layer
    us, err = run_layer(B=512, N=1024, K=1024, reps=3)
    assert err < 5e-6 and us > 0
