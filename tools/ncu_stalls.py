"""Summarise the per-instruction stall sampling of one kernel from `ncu -i X.ncu-rep --page source --csv`.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --launch-skip N --launch-count 1 > /tmp/src.csv
    python tools/ncu_stalls.py /tmp/src.csv [top_n]
"""
import csv
import sys

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
print(rows[0][:2])
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[h.index("# Samples")].isdigit() and r[0].startswith("0x")]
iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data), "warp-instructions executed", sum(int(r[iEx]) for r in data))
agg = {h[i]: sum(int(r[i]) for r in data) for i in stall_cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
print(f"--- top {top_n} instructions by samples (index, SASS, samples, executed, top stall reasons)")
order = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top_n]
for i in sorted(order):
    r = data[i]
    st = {h[j]: int(r[j]) for j in stall_cols if int(r[j]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(i, r[iSrc].strip()[:80], r[iS], r[iEx], top)
