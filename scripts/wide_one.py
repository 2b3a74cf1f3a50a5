"""Three fp16 passes of the wide policy over 18944 rows, or argv[1] rows (ncu target: 4 wide_gemm launches per pass)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_onnx_controller_b200 import capi
from go2_onnx_controller_b200.actor import PolicyBatch
from oracle import onnx_mini

ws, bs = onnx_mini.make_wide_policy(seed=5)
path = os.path.join(tempfile.mkdtemp(), "wide.onnx")
open(path, "wb").write(onnx_mini.write_mlp_onnx(ws, bs, 1.0, batch="batch"))
p = PolicyBatch(path, history=5)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
x = torch.randn(B, 245, device="cuda")
y = torch.empty(B, 12, device="cuda")
for _ in range(3):
    p.infer_device(x.data_ptr(), y.data_ptr(), B, capi.PREC_FP16)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
p.close()
