"""Per-source-line instruction counts and stall-sample shares of the first kernel of an .ncu-rep
(needs -lineinfo builds and --import-source on).  usage: ncu_hot_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--print-source', 'cuda,sass',
                      '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[2]
ix = {}
for j, k in enumerate(hdr):
    ix.setdefault(k, j)
agg, fname = {}, None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if len(r) > 10 and r[2] == '-' and r[0].isdigit():
        a = agg.setdefault((fname, int(r[0])), [0, 0, r[1][:84]])
        a[0] += int(r[ix['# Samples']])
        a[1] += int(r[ix['Instructions Executed']])
ts = sum(a[0] for a in agg.values())
te = sum(a[1] for a in agg.values())
print(f"warp instructions executed {te}, stall samples {ts}")
for (f, ln), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{f}:{ln:<5d} instr {100 * e / te:5.1f}%  samples {100 * s / ts:5.1f}%  {src}")
