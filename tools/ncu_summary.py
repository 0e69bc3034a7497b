"""Summarise an .ncu-rep (one kernel) into profiles/<name>.json + .md: the numbers bench.py / DESIGN.md cite.
    python tools/ncu_summary.py gpurun_out/prof_r1c.ncu-rep r1c_forward_v3 "<note>"
"""
import csv, io, json, os, subprocess, sys

rep, name = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__inst_executed_op_tma_ld.sum', 'sass__inst_executed_shared_loads',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg']
out = {k: {'value': m[k][0], 'unit': m[k][1]} for k in want if k in m}


def num(k):
    v, u = m[k]
    v = float(v.replace(',', ''))
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1}.get(u, 1)
    return v * scale


out['derived'] = {'dram_bytes_total': num('dram__bytes_read.sum') + num('dram__bytes_write.sum'),
                  'duration_s_under_ncu': num('gpu__time_duration.sum')}
out['note'] = note
out['source'] = os.path.basename(rep)
with open(os.path.join(ROOT, 'profiles', name + '.json'), 'w') as f:
    json.dump(out, f, indent=1)
with open(os.path.join(ROOT, 'profiles', name + '.md'), 'w') as f:
    f.write('# ncu summary: %s\n\n%s\n\nsource capture: `%s` (`ncu --set full --clock-control none --import-source on`)\n\n'
            '| metric | value | unit |\n|---|---|---|\n' % (name, note, os.path.basename(rep)))
    for k in want:
        if k in m:
            f.write('| %s | %s | %s |\n' % (k, m[k][0], m[k][1]))
    f.write('| dram bytes read+write (traffic) | %.0f | byte |\n' % out['derived']['dram_bytes_total'])
print(json.dumps(out['derived']))
