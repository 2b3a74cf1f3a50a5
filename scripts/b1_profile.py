"""Bounded twin of the batch-1 resident kernel for ncu (profiles/): N closed-loop steps in ONE launch, weights and
history in shared memory.  ncu's dram__bytes over the launch / N = HBM bytes per control step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import go2_onnx_controller_b200 as pkg
from go2_onnx_controller_b200 import capi
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
raws = list(bench.synthetic_raw_states(capi, 256, seed=2))
ctl = pkg.Go2Controller(pkg.DEFAULT_MODEL, b1_mode=capi.B1_LAUNCH)
for _ in range(2):
    act, ms = ctl.selfdriven(raws, steps)
print(f"selfdriven: {steps} steps in {ms:.3f} ms = {ms * 1e3 / steps:.3f} us/step; last action[0]={act[0]:.6f}")
ctl.close()
