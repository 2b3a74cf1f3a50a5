"""Time the wide-policy (245-1024-512-256-12) batched paths on one GPU: per-launch ms and inferences/s."""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from go2_onnx_controller_b200 import capi
from go2_onnx_controller_b200.actor import PolicyBatch
from oracle import onnx_mini

ws, bs = onnx_mini.make_wide_policy(seed=5)
path = os.path.join(tempfile.mkdtemp(), "wide.onnx")
open(path, "wb").write(onnx_mini.write_mlp_onnx(ws, bs, 1.0, batch="batch"))
p = PolicyBatch(path, history=5)
flops_row = 2 * sum(w.shape[0] * w.shape[1] for w in ws)
for B in (4096, 18944, 151552, 606208):
    x = torch.randn(B, 245, device="cuda")
    y = torch.empty(B, 12, device="cuda")
    for prec, name in ((capi.PREC_FP16, "fp16"), (capi.PREC_BF16, "bf16"), (capi.PREC_FP32, "fp32")):
        if prec == capi.PREC_FP32 and B > 151552:
            continue
        p.time_device(x.data_ptr(), y.data_ptr(), B, prec, 3)
        ms = p.time_device(x.data_ptr(), y.data_ptr(), B, prec, 10) / 10   # total ms over iters -> per launch
        print(f"B={B:7d} {name}: {ms:8.4f} ms  {B / ms * 1e3:.3e} inf/s  {B * flops_row / ms * 1e-9:.1f} TFLOP/s  launches={p.last_launches()}")
p.close()
