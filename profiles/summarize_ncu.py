"""Dev tool: condense an .ncu-rep (read here with `ncu -i`, no GPU needed) into the handful of counters the
roofline discussion in DESIGN.md uses.   python profiles/summarize_ncu.py gpurun_out/leaf_r1.ncu-rep > profiles/leaf_r1.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
STALLS = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit = rows[0], rows[1]
for r in rows[2:]:
    rec = dict(zip(hdr, r))
    print("kernel:", rec.get("Kernel Name"), " grid", rec.get("Grid Size"), " block", rec.get("Block Size"))
    for h, u, v in zip(hdr, unit, r):
        if h in WANT:
            print("  %-72s %14s %s" % (h, v, u))
    stalls = [(float(v), h[len(STALLS):-len("_per_issue_active.ratio")]) for h, v in zip(hdr, r)
              if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and v]
    print("  warp stall reasons (warps stalled per issue-active cycle):")
    for v, n in sorted(stalls, reverse=True)[:8]:
        print("    %-28s %.3f" % (n, v))
