"""Print the clock64 timeline an instrumented build (make EXTRA=-DHV_TC_INSTRUMENT, HIDVAE_TC_DEBUG=64) writes: python tools/ts_show.py LOG [n_launch]"""
import sys, collections
rows = [l.split() for l in open(sys.argv[1]) if l.startswith("TS")]
# launches are separated by timestamps going far backwards per 'who'; split into launches by 'who 0' restarts (code 1 after code>1)
launch, cur = [], []
for _, w, c, t in rows:
    w, c, t = int(w), int(c), int(t)
    if w == 0 and c == 1 and cur and any(x[0] == 1 for x in cur):
        launch.append(cur); cur = []
    cur.append((w, c, t))
launch.append(cur)
sel = int(sys.argv[2]) if len(sys.argv) > 2 else len(launch) - 1
ev = launch[sel]
print("launches", len(launch), "showing", sel, "events", len(ev))
for who in (0, 1):
    e = [(c, t) for w, c, t in ev if w == who]
    if not e: continue
    print("who", who, " ".join(f"{c}:+{(t - e[i-1][1]) & 0xffffffff if i else 0}" for i, (c, t) in enumerate(e)))
    d = collections.defaultdict(list)
    for i in range(1, len(e)):
        d[(e[i-1][0], e[i][0])].append((e[i][1] - e[i-1][1]) & 0xffffffff)
    for k, v in sorted(d.items()): print("   ", k, "n", len(v), "mean", round(sum(v) / len(v)), "min", min(v), "max", max(v))
    print("    span", (e[-1][1] - e[0][1]) & 0xffffffff)
