"""Copy the evidence of one gpurun round (tools/gpu_round.sh) from gpurun_out/ into profiles/ under a tag:
    python tools/make_profiles.py r01_v11
bench JSON lines, the ncu launch list, a per-launch summary of the `ncu --set full` capture, the DRAM traffic table that
bench.py echoes as roofline.traffic, and the per-instruction stall tables of the two long kernels."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(src, "bench.json"), os.path.join(dst, f"{tag}_bench.json"))
shutil.copy(os.path.join(src, "bench_ref.json"), os.path.join(dst, f"{tag}_bench_reference_arm.json"))
shutil.copy(os.path.join(src, "launches.csv"), os.path.join(dst, f"{tag}_launches.csv"))
rep = os.path.join(src, "prof_tc.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
keep = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
idx = [h.index(k) for k in keep if k in h]
with open(os.path.join(dst, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
name, rd, wr, ms = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
unit = rows[1][rd]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[unit]
def total(r): return int((float(r[rd]) + float(r[wr])) * scale)
launches = rows[2:]
small = {"train_forward": None, "backward": None, "eval_encode": None}
big = dict(small)
for r in launches:  # first three = the C2 step, the long ones = the 1 Mi-row launches
    nm = r[name]
    # rq_fwd_tc_v11_kernel<ROT, OUT>: OUT = 1 is the training forward (outputs), <0, 0> the ids-only encode
    kind = "backward" if "rq_bwd_kernel" in nm else (None if "rq_fwd_tc" not in nm else ("eval_encode" if ("<0, 0>" in nm or "(bool)0, (bool)0" in nm) else "train_forward"))
    if kind is None: continue
    tgt = big if float(r[ms]) > 100 else small
    if tgt[kind] is None: tgt[kind] = total(r)
n1, d, L, k = 1 << 20, 32, 3, 256
n2 = 12101
alg = lambda n: {"train_forward": n * (4 * d + 4 * d * L + 8 * L + 4), "backward": n * (4 * d + 8 * L + 4 * d * L + 4 + 4 * d) + 4 * L * k * d,
                 "eval_encode": n * (4 * d + 8 * L)}
json.dump({"source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none; short launches = the C2 step, long ones = 1 Mi rows)",
           "c2_bytes_per_launch": small, "c2_algorithmic_bytes": alg(n2), "rows_1Mi_bytes_per_launch": big, "rows_1Mi_algorithmic_bytes": alg(n1),
           "note": "C2 launches: writes stay in L2 (126 MB) within the capture, so dram write is ~0; reads = x + operand images / codebooks + gradients"},
          open(os.path.join(dst, "r01_traffic.json"), "w"), indent=1)
for kind, pat in (("encode_1Mi", "rq_fwd_tc"), ("backward_1Mi", "rq_bwd_kernel")):
    skip = next(i for i, r in enumerate(launches) if pat in r[name] and float(r[ms]) > 100)
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    tmp = f"/tmp/{tag}_{kind}.csv"
    open(tmp, "w").write(srcp)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), tmp, "40"], capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{tag}_ncu_stalls_{kind}.txt"), "w").write(out)
print(open(os.path.join(dst, "r01_traffic.json")).read())
