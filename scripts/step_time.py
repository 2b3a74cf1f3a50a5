"""Times the two launches of go2p_step_batch separately (1,048,576 robots)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi
import bench
rows = 1_048_576
pb = pkg.PolicyBatch(pkg.DEFAULT_MODEL)
arr = bench.synthetic_raw_states(capi, 4096, seed=3)
raw_np = np.frombuffer(bytes(arr), np.uint8).reshape(4096, C.sizeof(capi.RawState))
d_raw = torch.from_numpy(np.tile(raw_np, (rows // 4096, 1)).copy()).cuda()
obs = torch.zeros((rows, 98), device="cuda"); vel = torch.zeros((rows, 3), device="cuda")
act = torch.zeros((rows, 12), device="cuda"); q = torch.zeros((rows, 12), device="cuda", dtype=torch.float64)
b0 = torch.zeros(rows, device="cuda", dtype=torch.int32)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("assemble      %.4f ms" % timed(lambda: pb.assemble_device(d_raw.data_ptr(), act.data_ptr(), vel.data_ptr(), obs.data_ptr(), rows)))
print("infer plain   %.4f ms" % timed(lambda: pb.infer_device(obs.data_ptr(), act.data_ptr(), rows, capi.PREC_FP16)))
print("infer clamp   %.4f ms" % timed(lambda: pb.infer_device(obs.data_ptr(), act.data_ptr(), rows, capi.PREC_FP16, 0, b0.data_ptr(), None, capi.F_CLAMP_MASK)))
print("infer cl+qdes %.4f ms" % timed(lambda: pb.infer_device(obs.data_ptr(), act.data_ptr(), rows, capi.PREC_FP16, 0, b0.data_ptr(), q.data_ptr(), capi.F_CLAMP_MASK | capi.F_QDES)))
print("step_batch    %.4f ms" % timed(lambda: pb.step_device(d_raw.data_ptr(), vel.data_ptr(), obs.data_ptr(), act.data_ptr(), q.data_ptr(), rows, capi.PREC_FP16)))
