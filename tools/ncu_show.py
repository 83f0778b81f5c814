"""Key counters of every launch in an ncu report:  python tools/ncu_show.py X.ncu-rep [more_metric_substrings...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
extra = sys.argv[2:]
name = h.index("Kernel Name")
for r in rows[2:]:
    print("==", r[name][:100], "grid", r[h.index("Grid Size")], "block", r[h.index("Block Size")])
    for i, k in enumerate(h):
        base = k.split(".", 2)[-1] if k.count(".") > 2 and k.split(".")[1] in ("TriageCompute",) else k
        if k in KEYS or any(e in k for e in extra):
            print(f"   {k:95s} {r[i]:>18s} {u[i]}")
