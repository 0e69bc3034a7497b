"""The hi / lo plane layout of the experimental tensor-core path (csrc/forward_tc.cu `tc_plane_off`, csrc/probe_tc.cu) must
be the canonical K-major no-swizzle UMMA layout: checked on the host against CuTe's own
`tile_to_shape(UMMA::Layout_K_INTER_Atom<tfloat32_t>, [rows x 16])` from the CUTLASS headers vendored in this image,
together with the LBO / SBO the shared-memory descriptors are built with, and the instruction / shared-memory
descriptor bit patterns against `UMMA::make_instr_desc` / `UMMA::SmemDescriptor`."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
using namespace cute;
template <int ROWS>
int check() {
    auto layout = tile_to_shape(UMMA::Layout_K_INTER_Atom<tfloat32_t>{}, Shape<Int<ROWS>, _16>{});
    int bad = 0;
    for (int r = 0; r < ROWS; ++r)
        for (int k = 0; k < 16; ++k)
            bad += ((k >> 2) * ROWS * 4 + (r >> 3) * 32 + (r & 7) * 4 + (k & 3)) != (int)layout(r, k);
    // byte strides the descriptor carries: SBO = next 8-row group, LBO = next 16-byte K chunk
    printf("ROWS %d bad %d sbo %d lbo %d cosize %d\n", ROWS, bad, 4 * ((int)layout(8, 0) - (int)layout(0, 0)),
           4 * ((int)layout(0, 4) - (int)layout(0, 0)), (int)cosize(layout));
    return bad;
}
int main() {
    // descriptor encodings as CUTLASS builds them (cute/arch/mma_sm100_desc.hpp)
    auto d1 = UMMA::make_instr_desc<tfloat32_t, tfloat32_t, float, 128, 256, UMMA::Major::K, UMMA::Major::K>();
    auto d2 = UMMA::make_instr_desc<tfloat32_t, tfloat32_t, float, 128, 128, UMMA::Major::K, UMMA::Major::K>();
    UMMA::SmemDescriptor s;
    s.desc_ = 0;
    s.start_address_ = (0x12340 >> 4) & 0x3FFF, s.leading_byte_offset_ = 2048 >> 4, s.stride_byte_offset_ = 128 >> 4, s.version_ = 1;
    s.base_offset_ = 0, s.lbo_mode_ = 0, s.layout_type_ = 0;
    printf("idesc256 %u idesc128 %u sdesc %llu\n", (uint32_t)d1.desc_, (uint32_t)d2.desc_, (unsigned long long)s.desc_);
    return check<128>() + check<256>();
}
'''


def _cutlass_include():
    import importlib.util
    spec = importlib.util.find_spec('flashinfer')   # located, not imported
    if spec is None or not spec.submodule_search_locations:
        return None
    path = os.path.join(list(spec.submodule_search_locations)[0], 'data', 'cutlass', 'include')
    return path if os.path.exists(os.path.join(path, 'cute', 'tensor.hpp')) else None


def test_plane_layout_is_cute_canonical_k_major(tmp_path):
    inc = _cutlass_include()
    if inc is None:
        pytest.skip("CUTLASS / CuTe headers not found in this image")
    src = tmp_path / 'chk.cpp'
    src.write_text(SRC)
    exe = tmp_path / 'chk'
    subprocess.run(['g++', '-std=c++17', '-I', inc, '-I', '/usr/local/cuda/include', str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    got = {int(m.group(1)): tuple(int(v) for v in m.groups()[1:])
           for m in re.finditer(r'ROWS (\d+) bad (\d+) sbo (\d+) lbo (\d+) cosize (\d+)', out.stdout)}
    # (mismatches, SBO bytes, LBO bytes, floats per plane): forward_tc.cu TC_SBO / TC_LBO_A / TC_LBO_W, probe_tc.cu P_SBO / P_LBO
    assert got[128] == (0, 128, 128 * 16, 128 * 16) and got[256] == (0, 128, 256 * 16, 256 * 16)
    # instruction / shared-memory descriptor bit patterns: the constants of forward_tc.cu / probe_tc.cu, re-stated
    idesc = lambda m, n: (1 << 4) | (2 << 7) | (2 << 10) | ((n >> 3) << 17) | ((m >> 4) << 24)  # noqa: E731
    sdesc = ((0x12340 & 0x3FFFF) >> 4) | (((2048 >> 4) & 0x3FFF) << 16) | (((128 >> 4) & 0x3FFF) << 32) | (1 << 46)
    m = re.search(r'idesc256 (\d+) idesc128 (\d+) sdesc (\d+)', out.stdout)
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (idesc(128, 256), idesc(128, 128), sdesc)
    cu = open(os.path.join(ROOT, 'monoloco_b200', 'csrc', 'forward_tc.cu')).read()
    assert '(1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24)' in cu
    assert 'TC_LBO_A = TCM * 16, TC_LBO_W = TCN * 16, TC_SBO = 128' in cu
    assert '(size_t)(k_in_block >> 2) * tile_rows * 4 + (size_t)(r >> 3) * 32 + (size_t)(r & 7) * 4 + (k_in_block & 3)' in cu


def test_tf32x3_emulation_meets_parity_rule():
    """The arithmetic the experimental tensor-core path implements (tools/tf32x3_study.py: operands split into two TF32
    terms, cross terms in their own accumulator, fp32 accumulators rounded once per k = 8 MMA -- even with the pessimistic
    truncating adder) keeps the whole LocoModel forward inside the product's parity rule; plain TF32 does not."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from monoloco_b200 import synthetic
    from oracle import loco_oracle as O
    from tools.tf32x3_study import forward, mm_tf32x3_split, to_tf32
    sd = {k: np.asarray(v) for k, v in synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0).items()}
    x = O.preprocess_monoloco(synthetic.make_keypoints(24, seed=0), synthetic.KITTI_K)
    ref = O.loco_model_forward(sd, x)
    out = forward(sd, x, lambda a, w: mm_tf32x3_split(np.asarray(a, np.float32), w, 'rz'))
    ok, worst = O.close(out, ref)
    assert ok and worst < 1.0, worst
    plain = forward(sd, x, lambda a, w: (to_tf32(np.asarray(a, np.float32)).astype(np.float64) @ to_tf32(w).astype(np.float64).T).astype(np.float32))
    assert not O.close(plain, ref)[0]
