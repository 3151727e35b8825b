"""Turns the raw ncu outputs of a gpurun call into the summaries committed under profiles/.
Usage: python scripts/summarise_profiles.py <round-tag>   (reads gpurun_out/launches.csv and
gpurun_out/prof_bench.ncu-rep; needs ncu on PATH, no GPU)."""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

# ---- launch list -> per-kernel shares of one bench step
rows = []
with open("gpurun_out/launches.csv") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((int(r["ID"]), r["Kernel Name"], v))
open(f"profiles/{tag}_bench_launches.csv", "w").write("".join(lines))
# one steady-state bench step = [som_apply, codebook_prep, whole-pass kernel (ACC variant, "..., 1>"),
# codebook_prep, assign kernel ("..., 0>"), bmu_exact fix-up]; take the last such window of full-size
# launches (the oracle check on a sampled shard and the chunked end-to-end leg follow in the list)
pat = ["som_apply", "codebook_prep", ", 1>", "codebook_prep", ", 0>", "bmu_exact"]
wins = [i for i in range(len(rows) - 5)
        if all(pat[j] in rows[i + j][1] for j in range(6)) and rows[i + 2][2] > 1.0 and rows[i + 4][2] > 1.0]
i = wins[-1]
step = rows[i:i + 6]
tot = sum(v for _, _, v in step)
with open(f"profiles/{tag}_bench_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled "
            "-k regex:5pixie -c 80: python bench.py --steps 2 --warmup 3 --no-cfg3 (N=1)\n"
            "Only this library's kernels are listed (the synthetic-data generation is torch and is "
            "filtered out).  The list holds 5 bench steps (3 warm-up + 2 timed), the bench's exact-kernel "
            "spot check (one 3.1 ms bmu_exact launch) and the chunked launches of the end-to-end leg.\n"
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
            f"One steady-state step (launch IDs {step[0][0]}-{step[-1][0]}), {tot*1e3:.1f} us of kernel time:\n")
    for _, k, v in step:
        f.write(f"{v*1e3:10.1f} us  {100*v/tot:5.1f} %  {k.split('(')[0]}\n")
    f.write("\nbench.py (CUDA events, warm) for the same step: see the bench line "
            "(train_ms_per_step / assign_ms_per_step).\n")
print(open(f"profiles/{tag}_bench_launches_summary.txt").read())

# ---- full capture -> selected metrics per kernel
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_bench.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rd[0], rd[1], rd[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = ["ncu --set full --clock-control none --import-source on -k regex:bmu_tc_kernel -s 4 -c 2 "
       "python bench.py --steps 3 --warmup 1  (N=1, BASELINE cfg2: 52,428,800 rows x 32 channels, "
       "10x10 SOM, Pixie-like rows, codebook trained by the bench)",
       "Two consecutive launches of one bench step: the whole-pass training kernel (<..., 1> = ACC) "
       "and the assign kernel (<..., 0>).",
       "Times under ncu are cold-cache and serialised; bench.py times the same launches with CUDA events.",
       ""]
traffic = None
for row in data:
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    out.append(f"{'Kernel Name':95s} {d['Kernel Name']}")
    for m in want:
        if m in d:
            out.append(f"{m:95s} {d[m]} {u[m]}")
    out.append("")
    if d["Kernel Name"].rstrip().endswith("0>(CUtensorMap_st, TcParams)"):
        def to_bytes(m):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[m]]
            return float(d[m].replace(",", "")) * scale
        traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
open(f"profiles/{tag}_bench_ncu_summary.txt", "w").write("\n".join(out))
print("\n".join(out[:40]))
if traffic:
    json.dump({"dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": 52428800 * 132,
               "source": f"profiles/{tag}_bench_ncu_summary.txt (ncu --set full, assign launch of one "
                         "bench step)"}, open(f"profiles/{tag}_assign_traffic.json", "w"), indent=1)
    print("traffic", traffic)
