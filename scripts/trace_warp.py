import sys, numpy as np
t=np.load('gpurun_out/tc_trace_raw.npy')
names = {1: "P issue", 2: "M ready", 3: "M commit", 4: "W obs_full", 5: "W conv done", 6: "W E start", 7: "W E done", 8: "W out start", 9: "W out done", 10: "W computed", 11: "W prefetched", 12: "W st done", 13: "M probe hit (layer=rdy mask)", 14: "M issued", 15: "W fine (layer=substep)"}
w=int(sys.argv[1]); lo=int(sys.argv[2]); cnt=int(sys.argv[3]) if len(sys.argv)>3 else 30
n=int(t[w,2046]); ev=t[w,0:2*n:2]; ck=t[w,1:2*n:2]
c0=t[:18,1].min()
prev=None
for e,c in list(zip(ev,ck))[lo:lo+cnt]:
    e=int(e); k=e>>8; l=(e>>4)&15; s=e&1
    print(f"{c-c0:9d} (+{(c-prev) if prev else 0:5d}) {names.get(k,hex(e)):12s} slot {s} layer {l}")
    prev=c
