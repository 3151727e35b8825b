"""Key raw metrics of the first kernel in an .ncu-rep (run where ncu is installed).
usage: ncu_key_metrics.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
ix = {n: i for i, n in enumerate(h)}
print(v[ix['Kernel Name']])
for w in WANT:
    if w in ix:
        print(f"{w:75s} {v[ix[w]]} {units[ix[w]]}")
for n, i in ix.items():
    if 'issue_stalled' in n and n.endswith('per_issue_active.ratio') and 'not_issued' not in n:
        try:
            if float(v[i]) >= 0.3:
                print(f"{n:75s} {float(v[i]):.2f}")
        except ValueError:
            pass
for n, i in ix.items():
    if 'tensor' in n and 'pct' in n:
        print(f"{n:75s} {v[i]}")
