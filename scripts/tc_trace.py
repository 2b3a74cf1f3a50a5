"""Timeline of one CTA of tc_mlp_kernel (trace build: -DGO2P_TC_TRACE).  Prints per-event clock deltas."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GO2P_LIB", os.path.join(ROOT, "go2_onnx_controller_b200", "lib", "libgo2policy_trace.so"))
import numpy as np, torch
trace = torch.zeros(34 * 2048, dtype=torch.int64, device="cuda")
os.environ["GO2P_TC_TRACE_PTR"] = str(trace.data_ptr())
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_048_576
pb = pkg.PolicyBatch(pkg.DEFAULT_MODEL)
d_obs = torch.randn((B, 98), device="cuda"); d_act = torch.empty((B, 12), device="cuda")
for _ in range(3):
    trace.zero_()
    torch.cuda.synchronize()
    pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), B, capi.PREC_FP16)
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(34, 2048)
evs, cks = [], []
for w in range(34):
    n = int(t[w, 2046])
    evs.append(t[w, 0:2 * n:2]); cks.append(t[w, 1:2 * n:2])
np.save(os.path.join(ROOT, "gpurun_out", "tc_trace_raw.npy"), t)
ev = np.concatenate(evs); ck = np.concatenate(cks); n = ev.size
order = np.argsort(ck, kind="stable"); ev = ev[order]; ck = ck[order] - ck[order][0]
names = {1: "P issue", 2: "M ready", 3: "M commit", 4: "E obs_full", 5: "E L0 done", 6: "E acc_full", 7: "E layer done", 8: "E out start", 9: "E out done"}
np.save(os.path.join(ROOT, "gpurun_out", "tc_trace.npy"), np.stack([ev, ck]))
print("events", n, "span cycles", ck[-1])
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 400
for e, c in list(zip(ev, ck))[lo:lo + 75]:
    k = int(e) >> 8; l = (int(e) >> 4) & 7; s = int(e) & 1; w7 = (int(e) >> 3) & 1
    print(f"{c:9d}  {names.get(k, hex(e)):14s} slot {s} layer {l} {'(wq7)' if w7 else ''}")
