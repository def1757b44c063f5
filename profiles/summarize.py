#!/usr/bin/env python
"""Summarise an ncu report + launch list into a small text file under profiles/.
Usage: python profiles/summarize.py gpurun_out/prof_TAG.ncu-rep gpurun_out/launches_TAG.csv profiles/TAG_summary.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, launches, out = sys.argv[1:4]
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum.per_second", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
    for r in rows[2:]:
        f.write(f"\n## {r[idx['Kernel Name']]}\n")
        for m in METRICS:
            if m in idx:
                f.write(f"{m:75s} {r[idx[m]]} {units[idx[m]]}\n")
    f.write(f"\n# launch list ({launches}): gpu__time_duration.sum per launch, ns (cold-cache, serialised)\n")
    per = collections.defaultdict(list)
    for r in csv.reader(open(launches)):
        if len(r) > 5 and r[0].isdigit():
            per[r[4].split("(")[0]].append(float(r[-1]))
    tot = sum(sum(v) for v in per.values())
    for k, v in per.items():
        f.write(f"{k:60s} launches {len(v):3d}  mean {sum(v)/len(v)/1e6:8.3f} ms  share {100*sum(v)/tot:5.1f} %\n")
print(open(out).read())
