"""SASS opcode summary of every kernel in libmonoloco_b200.so (cuobjdump -sass): the mnemonics that prove which hardware
paths a kernel uses (B200_PROFILING.md: UTC*MMA = tcgen05.mma, UBLKCP / UTMALDG = TMA, LDTM / STTM = Tensor Memory,
FFMA2 = packed fp32 FMA).   python tools/sass_summary.py [out.md]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'monoloco_b200', 'lib', 'libmonoloco_b200.so')
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r2_sass_summary.md')
txt = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
WATCH = ['FFMA2', 'FFMA', 'FMUL2', 'FADD2', 'UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'HMMA', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM',
         'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'ATOMG', 'RED', 'BAR', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'MEMBAR', 'DFMA', 'MUFU',
         'SHFL', 'ST', 'LD']
kern = None
counts = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        counts[kern][op] += 1
        counts[kern]['_total'] += 1
with open(out_path, 'w') as f:
    f.write('# SASS opcode summary of libmonoloco_b200.so (sm_100a)\n\n`cuobjdump -sass` of the in-tree library, instruction counts per kernel '
            '(static, one count per SASS line).  `UTC*MMA` = tcgen05.mma, `UBLKCP` = 1-D TMA bulk copy, `LDTM` / `STTM` = Tensor Memory load / store, '
            '`FFMA2` = packed fp32 FMA, `UCGABAR_*` = cluster barrier.\n\n')
    cols = ['_total', 'FFMA2', 'FFMA', 'UTCHMMA', 'UBLKCP', 'LDTM', 'STTM', 'LDS', 'STS', 'LDG', 'STG', 'ST', 'SYNCS', 'UTCBAR', 'BAR', 'UCGABAR_ARV', 'ATOMG', 'RED', 'DFMA', 'HMMA']
    f.write('| kernel | ' + ' | '.join(c.strip('_') for c in cols) + ' |\n|---|' + '---|' * len(cols) + '\n')
    for k, c in counts.items():
        if c['_total'] < 40:
            continue
        f.write('| `%s` | ' % k + ' | '.join(str(c.get(col, 0)) for col in cols) + ' |\n')
print(out_path, len(counts), 'kernels')
