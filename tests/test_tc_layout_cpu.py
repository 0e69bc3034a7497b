"""The hi / lo plane layout of the experimental tensor-core path (csrc/forward_tc.cu `tc_plane_off`, csrc/probe_tc.cu) must
be the canonical K-major no-swizzle UMMA layout: checked on the host against CuTe's own
`tile_to_shape(UMMA::Layout_K_INTER_Atom<tfloat32_t>, [rows x 16])` from the CUTLASS headers vendored in this image,
together with the LBO / SBO the shared-memory descriptors are built with."""
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
int main() { return check<128>() + check<256>(); }
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
    cu = open(os.path.join(ROOT, 'monoloco_b200', 'csrc', 'forward_tc.cu')).read()
    assert 'TC_LBO_A = TCM * 16, TC_LBO_W = TCN * 16, TC_SBO = 128' in cu
    assert '(size_t)(k_in_block >> 2) * tile_rows * 4 + (size_t)(r >> 3) * 32 + (size_t)(r & 7) * 4 + (k_in_block & 3)' in cu
