"""Where the pool's time goes, from a tc_trace.py capture: per warp, time between consecutive trace points grouped by
(from-event, to-event) kind."""
import numpy as np, sys, collections
t = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/tc_trace_raw.npy')
names = {1: "tma", 2: "a_ready", 3: "committed", 4: "conv acq", 6: "E acq", 7: "arrived", 8: "out acq", 9: "out done", 10: "A free",
         11: "st drained", 13: "acq begin", 14: "st issued", 15: "fine"}
nw = max(w for w in range(34) if t[w, 2046] > 0)
for w in list(range(0, 16)) + [nw]:
    n = int(t[w, 2046]); ev = [int(e) for e in t[w, 0:2 * n:2]]; ck = [int(c) for c in t[w, 1:2 * n:2]]
    lo = n // 4
    agg = collections.OrderedDict()
    for i in range(lo, n - 1):
        a, b = ev[i], ev[i + 1]
        ka = names.get(a >> 8, hex(a >> 8)) + (str((a >> 4) & 15) if (a >> 8) == 15 else "")
        kb = names.get(b >> 8, hex(b >> 8)) + (str((b >> 4) & 15) if (b >> 8) == 15 else "")
        agg.setdefault((ka, kb), []).append(ck[i + 1] - ck[i])
    tot = ck[n - 1] - ck[lo]
    print(f"--- warp {w} (q{w % 4} cb{w // 4}): {n} events, {tot} cycles after event {lo}")
    for (ka, kb), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) > tot * 0.02:
            print(f"   {ka:>12s} -> {kb:<12s} n {len(v):4d}  median {int(np.median(v)):6d}  total {100 * sum(v) / tot:5.1f} %")
