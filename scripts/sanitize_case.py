"""Small invocations of the three kernel families for compute-sanitizer (scripts/gpu_sanitize.sh): the batched tcgen05
kernel (mbarrier / TMEM / bulk-copy pipeline), the wide per-layer tcgen05 GEMMs and the bounded twin of the resident
batch-1 kernel, each checked against the oracle so a sanitizer-clean run is also a correct one."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi, onnx_writer
from oracle import oracle
import bench

pol = oracle.load_policy(pkg.DEFAULT_MODEL)
pb = pkg.PolicyBatch(pkg.DEFAULT_MODEL)
for B in (1, 300, 148 * 128 * 2 + 77):                     # one ragged tile, a few tiles, > 2 tiles per CTA + ragged tail
    X = oracle.make_obs_d1(B, 98, seed=B)
    d_obs = torch.from_numpy(X).cuda(); d_act = torch.zeros((B, 12), device="cuda")
    d_q = torch.zeros((B, 12), device="cuda", dtype=torch.float64); d_b = torch.zeros(B, device="cuda", dtype=torch.int32)
    for prec in (capi.PREC_FP16, capi.PREC_BF16):
        pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), B, prec, 0, d_b.data_ptr(), d_q.data_ptr(), capi.F_CLAMP_MASK | capi.F_QDES)
        torch.cuda.synchronize()
        err = np.abs(d_act.cpu().numpy() - oracle.forward(pol, X)).max()
        assert err < 0.1, err
    print(f"tc_mlp_kernel B={B} ok")
pb.close()
wpath = os.path.join(tempfile.mkdtemp(), "wide.onnx")
onnx_writer.write_policy(wpath, *onnx_writer.wide_policy(5))
wp = pkg.PolicyBatch(wpath)
wpol = oracle.load_policy(wpath)
for B in (200, 19000):
    X = oracle.make_obs_d1(B, 245, seed=B)
    d_obs = torch.from_numpy(X).cuda(); d_act = torch.zeros((B, 12), device="cuda")
    wp.infer_device(d_obs.data_ptr(), d_act.data_ptr(), B, capi.PREC_FP16)
    torch.cuda.synchronize()
    assert np.abs(d_act.cpu().numpy() - oracle.forward(wpol, X)).max() < 5e-3
    print(f"wide_gemm_kernel B={B} ok")
wp.close()
ctl = pkg.Go2Controller(pkg.DEFAULT_MODEL, b1_mode=capi.B1_LAUNCH)
act, ms = ctl.selfdriven(list(bench.synthetic_raw_states(capi, 16, seed=2)), 50)
assert np.isfinite(act).all()
print("b1_selfdriven_kernel ok")
ctl.close()
