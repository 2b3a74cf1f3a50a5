import sys, numpy as np
t=np.load('gpurun_out/tc_trace.npy'); ev=t[0]; ck=t[1]
names = {1: "P issue", 2: "M ready", 3: "M commit", 4: "W obs_full", 5: "W conv done", 6: "W E start", 7: "W E done", 8: "W out start", 9: "W out done"}
t0=int(sys.argv[1]) if len(sys.argv)>1 else 150000; t1=t0+int(sys.argv[2]) if len(sys.argv)>2 else t0+16000
sel=(ck>=t0)&(ck<t1)
from collections import OrderedDict
groups=OrderedDict()
for e,c in zip(ev[sel],ck[sel]):
    groups.setdefault(int(e),[]).append(int(c))
# split groups by time gaps > 5000 (different pair)
rows=[]
for e,cs in groups.items():
    cs=sorted(cs); cur=[cs[0]]
    for c in cs[1:]:
        if c-cur[-1]>4000: rows.append((e,cur)); cur=[c]
        else: cur.append(c)
    rows.append((e,cur))
rows.sort(key=lambda r:r[1][0])
for e,cs in rows:
    k=e>>8; l=(e>>4)&7; s=e&1
    print(f"{cs[0]:8d} .. {cs[-1]:8d} (spread {cs[-1]-cs[0]:5d}) median {int(np.median(cs)):8d}  {names.get(k,hex(e)):12s} slot {s} layer {l} x{len(cs)}")
