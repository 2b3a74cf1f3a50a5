"""Times the batched tensor-core kernel (1,048,576 rows, fp16/bf16, plain and with the clamp/mask epilogue) and prints
its max abs error vs the fp64 oracle on a sample.  GO2P_LIB selects the library build under test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi
from oracle import oracle
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_048_576
model, width = pkg.DEFAULT_MODEL, 98
if os.environ.get("GO2P_TIME_IN"):      # synthetic Go2-topology policy with another input width
    import tempfile
    from go2_onnx_controller_b200 import onnx_writer
    width = int(os.environ["GO2P_TIME_IN"])
    ws, bs = onnx_writer.wide_policy(seed=3, dims=(width, 128, 128, 128, 12))
    model = os.path.join(tempfile.mkdtemp(), "narrow.onnx")
    onnx_writer.write_policy(model, ws, bs)
pb = pkg.PolicyBatch(model)
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn((rows, width), device="cuda", generator=g)
act = torch.zeros((rows, 12), device="cuda")
b0 = torch.zeros(rows, device="cuda", dtype=torch.int32)
pol = oracle.load_policy(model)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, prec in (("fp16", capi.PREC_FP16), ("bf16", capi.PREC_BF16)):
    t0 = timed(lambda: pb.infer_device(obs.data_ptr(), act.data_ptr(), rows, prec))
    t1 = timed(lambda: pb.infer_device(obs.data_ptr(), act.data_ptr(), rows, prec, 0, b0.data_ptr(), None, capi.F_CLAMP_MASK))
    idx = torch.randint(0, rows, (512,), device="cuda")
    ref = oracle.forward(pol, obs[idx].cpu().numpy())
    err = float(np.abs(act[idx].cpu().numpy() - ref).max())
    print(f"{os.path.basename(os.environ.get('GO2P_LIB', 'libgo2policy.so'))} {name}: plain {t0:.4f} ms  clamp {t1:.4f} ms  "
          f"frac {(4 * width + 48) * rows / (t1 * 1e-3) / 6455.6e9:.3f}  err {err:.2e}")
