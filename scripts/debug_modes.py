import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi
from oracle import oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import raw_struct, raw_py
g = np.load("tests/golden/go2_closed_loop_golden.npz")
pol = oracle.load_policy(pkg.DEFAULT_MODEL)
for mode in (0, 1, 2):
    ctl = pkg.Go2Controller(pkg.DEFAULT_MODEL, b1_mode=mode)
    st = oracle.ControllerState(H=2)
    for i in range(3):
        out = ctl.step(raw_struct(g, i))
        d_obs = np.frombuffer(out.observation, np.float32, 98).copy()
        d_raw = np.frombuffer(out.action_raw, np.float32, 12).copy()
        so = oracle.controller_step(pol, st, raw_py(g, i), np.float64, act_fn=lambda o: d_raw)
        bad = np.nonzero(d_obs != so.obs)[0]
        ref = oracle.forward(pol, d_obs)
        print(f"mode {mode} step {i}: obs mismatches at {bad.tolist()[:20]} (n={bad.size}); fwd err on device obs {np.abs(d_raw-ref).max():.3e}; "
              f"golden err {np.abs(d_raw-g['action_raw'][i]).max():.3e}")
        if i == 0 and mode != 0:
            print("   d_raw:", np.round(d_raw, 4).tolist())
            print("   ref  :", np.round(ref, 4).tolist())
            print("   d_act:", np.round(np.frombuffer(out.action, np.float32, 12), 4).tolist())
        if bad.size:
            print("   dev:", d_obs[bad][:12], "\n   ref:", so.obs[bad][:12])
    ctl.close()
