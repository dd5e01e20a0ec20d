"""Summarise an ncu report (raw + source pages): headline metrics and the top stall sites.
    python profiles/micro/ncu_top.py report.ncu-rep [ntop]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread']
for i, h in enumerate(hdr):
    if any(h == w or h.startswith(w + ".") and 'per_second' not in h for w in want) or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h):
        print(h, '=', vals[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, '# Samples') for r in data)
print('total samples', tot, 'instr rows', len(data))
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:ntop]:
    print(r[ix['Address']][-5:], '%-55s' % r[ix['Source']][:55], int(f(r, '# Samples')), 'long', int(f(r, 'stall_long_sb')),
          'short', int(f(r, 'stall_short_sb')), 'bar', int(f(r, 'stall_barrier')), 'wait', int(f(r, 'stall_wait')),
          'math', int(f(r, 'stall_math')), 'exec', int(f(r, 'Instructions Executed')))
