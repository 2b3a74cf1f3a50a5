"""Phase durations of the pool's E jobs from a tc_trace.py capture (gpurun_out/tc_trace_raw.npy)."""
import numpy as np, sys
t = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/tc_trace_raw.npy')
nw = max(w for w in range(34) if t[w, 2046] > 0)   # control warp index
c0 = min(int(t[w, 1]) for w in range(nw + 1) if t[w, 2046] > 0)
def events(w):
    n = int(t[w, 2046]); return [int(e) for e in t[w, 0:2 * n:2]], [int(c) - c0 for c in t[w, 1:2 * n:2]]
names = {1: "tma", 2: "a_ready", 3: "committed", 4: "conv acq", 6: "E acq", 7: "arrived", 8: "out acq", 9: "out done", 10: "A free", 11: "st drained", 15: "fine"}
for w in (0, 5, nw):
    ev, ck = events(w)
    print(f"--- warp {w}: {len(ev)} events, span {ck[-1] - ck[0]}")
    lo = len(ev) // 2
    prev = None
    for e, c in list(zip(ev, ck))[lo:lo + 45]:
        print(f"{c:9d} (+{(c - prev) if prev else 0:5d}) {names.get(e >> 8, hex(e)):12s} slot {e & 3} sub {(e >> 4) & 15}")
        prev = c
# job-level: time between consecutive 'arrived'/'out done' events of warp 0 = job durations
ev, ck = events(0)
ends = [(c, e) for e, c in zip(ev, ck) if (e >> 8) in (7, 9)]
d = np.diff([c for c, _ in ends])
print("warp 0 job durations: median", np.median(d), "mean", d.mean(), "p90", np.percentile(d, 90), "n", len(d))
kinds = {}
for (c1, e1), (c0_, e0) in zip(ends[1:], ends[:-1]):
    kinds.setdefault(e1 >> 8, []).append(c1 - c0_)
for k, v in kinds.items(): print(names[k], "median", np.median(v), "n", len(v))
