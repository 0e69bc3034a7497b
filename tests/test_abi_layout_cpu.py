"""C-ABI layout: the ctypes mirrors in monoloco_b200/_lib.py must match include/monoloco_b200.h field by field.
A small C program (gcc, plain C -- the header must stay C-clean) prints sizeof / offsetof of every struct member."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STRUCTS = {'mlb_op': 'MlbOp', 'mlb_model_desc': 'MlbModelDesc', 'mlb_forward_args': 'MlbForwardArgs',
           'mlb_train_block': 'MlbTrainBlock', 'mlb_train_args': 'MlbTrainArgs', 'mlb_post_args': 'MlbPostArgs'}


def _c_layout(tmp_path):
    from monoloco_b200 import _lib as L_
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "monoloco_b200.h"', 'int main(void) {']
    for cname, pyname in STRUCTS.items():
        cls = getattr(L_, pyname)
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  printf("abi %d ipc %d gather_ld %d gather_dec %d max_ops %d\\n", MLB_ABI_VERSION, MLB_IPC_HANDLE_BYTES,',
              '         MLB_GATHER_LD, MLB_GATHER_DEC, MLB_MAX_OPS);', '  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    return subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout


def test_ctypes_structs_match_header(tmp_path):
    from monoloco_b200 import _lib as L_
    out = _c_layout(tmp_path)
    got = {}
    for line in out.splitlines():
        parts = line.split()
        if parts[0] in STRUCTS:
            got[(parts[0], parts[1])] = int(parts[2])
    for cname, pyname in STRUCTS.items():
        cls = getattr(L_, pyname)
        assert got[(cname, 'sizeof')] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    consts = dict(zip(('abi', 'ipc', 'gather_ld', 'gather_dec', 'max_ops'),
                      map(int, re.search(r'abi (\d+) ipc (\d+) gather_ld (\d+) gather_dec (\d+) max_ops (\d+)', out).groups())))
    assert consts['abi'] == L_.MLB_ABI_VERSION and consts['ipc'] == L_.IPC_HANDLE_BYTES
    assert consts['gather_ld'] == L_.GATHER_LD and consts['gather_dec'] == L_.GATHER_DEC


def test_header_flag_values_match_python():
    from monoloco_b200 import _lib as L_
    hdr = open(os.path.join(ROOT, 'include', 'monoloco_b200.h')).read()
    for name, val in (('MLB_FWD_ZERO_CENTER', L_.FWD_ZERO_CENTER), ('MLB_FWD_DROPOUT', L_.FWD_DROPOUT),
                      ('MLB_FWD_RES_TMEM', L_.FWD_RES_TMEM), ('MLB_FWD_FORCE_TILE', L_.FWD_FORCE_TILE),
                      ('MLB_FWD_FORCE_CLUSTER', L_.FWD_FORCE_CLUSTER), ('MLB_FWD_RES_SCRATCH', L_.FWD_RES_SCRATCH),
                      ('MLB_FWD_FORCE_WIDE', L_.FWD_FORCE_WIDE), ('MLB_FWD_FORCE_TC', L_.FWD_FORCE_TC), ('MLB_FWD_FORCE_WIDE2', L_.FWD_FORCE_WIDE2)):
        m = re.search(name + r'\s*=\s*(\d+)', hdr)
        assert m and int(m.group(1)) == val, name
