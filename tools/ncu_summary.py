#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics per launch + top stall sites.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] [n_top]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'l1tex__lsu_writeback_active.sum',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:60], r[hdr.index('Grid Size')])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'   {w:75s} {r[i]:>16s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--kernel-name', 'regex:' + kern] if kern else []) +
                     ['--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[hi]
si, so, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print('total samples', tot)
agg = {}
for r in data:
    for i in stall:
        if r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print('stall mix:', sorted(((k[6:], round(100 * v / tot, 1)) for k, v in agg.items() if v), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[si]))[:ntop]:
    st = sorted(((hdr[i][6:], int(r[i])) for i in stall if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% ex={r[ie]:>9s} {r[so].strip()[:64]:64s} {st}")
