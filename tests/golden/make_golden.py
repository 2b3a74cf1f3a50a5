"""Generates the committed golden fixtures (run in the build container: python tests/golden/make_golden.py).

The reference holds no golden vector for this path and its arithmetic lives in the absent
onnxruntime 1.20.1 binary, so the vectors are produced by the numpy fp64 restatement (oracle/oracle.py) and
accepted only if an INDEPENDENT torch-CPU fp64 evaluation (F.linear / F.elu on weights parsed by a different
code path: the C reader in oracle/oracle_mlp.c) agrees to 1e-12.  The model file read here is the
reference's own onnx_inference/data/model.onnx when /root/reference exists (and must be byte-identical to the
bundled copy).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import coracle, oracle  # noqa: E402

BUNDLED = os.path.join(ROOT, "go2_onnx_controller_b200", "data", "model.onnx")
REF = "/root/reference/onnx_inference/data/model.onnx"
SHA = "9bdcb0f47be89417dbf3f7fb46522ecb3ec42ca65458a905f191aff933a18eae"


def torch_forward_f64(cm: coracle.CModel, X: np.ndarray) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    m = cm._p.contents
    x = torch.from_numpy(np.asarray(X, np.float32)).double()
    for l in range(m.n_layers):
        out_d, in_d = m.dims[l + 1], m.dims[l]
        w = torch.from_numpy(np.ctypeslib.as_array(m.w[l], shape=(out_d, in_d)).copy()).double()
        b = torch.from_numpy(np.ctypeslib.as_array(m.b[l], shape=(out_d,)).copy()).double()
        x = F.linear(x, w, b)
        if m.has_elu[l]:
            x = F.elu(x, alpha=float(m.elu_alpha[l]))
    return x.numpy()


def main():
    data = open(BUNDLED, "rb").read()
    assert hashlib.sha256(data).hexdigest() == SHA, "bundled model.onnx differs from the surveyed file"
    if os.path.exists(REF):
        assert open(REF, "rb").read() == data, "bundled model.onnx is not the reference's file"
    pol = oracle.load_policy(BUNDLED)
    coracle.build()
    cm = coracle.CModel(BUNDLED)

    kat_in = np.stack([np.zeros(98, np.float32), 2 * np.ones(98, np.float32)])   # main.cpp:32, main.py:20
    d1 = oracle.make_obs_d1(65536, 98, seed=0)[:512]
    d2 = oracle.make_obs_d2(pol, 256, seed=1)
    rng = np.random.default_rng(3)
    d3 = (d2 * np.float32(3000)).astype(np.float32)
    d3[5, :] = np.nan
    d3[6, 7] = np.inf
    d3[7, 11] = -np.inf
    d3[8, :] = -0.0
    d3[9, 40] = np.nan
    d3_button0 = (rng.random(256) < 0.05).astype(np.int32)
    d3_button0[[0, 5, 6]] = [1, 1, 0]

    out = {}
    for name, X in (("kat", kat_in), ("d1", d1), ("d2", d2), ("d3", d3)):
        with np.errstate(all="ignore"):
            y = oracle.forward(pol, X, np.float64)
            yt = torch_forward_f64(cm, X)
        fin = np.isfinite(y) & np.isfinite(yt)
        assert (np.isnan(y) == np.isnan(yt)).all(), name
        assert np.array_equal(y[~fin & ~np.isnan(y)], yt[~fin & ~np.isnan(yt)]), name
        err = np.abs(y[fin] - yt[fin]) / np.maximum(1.0, np.abs(y[fin]))
        assert err.max() < 1e-12, (name, err.max())
        out[f"{name}_obs"] = X
        out[f"{name}_action_f64"] = y
        print(f"{name}: {X.shape} torch-fp64 agreement {err.max():.2e}")
    # A9/A11 on the stress set: published action and joint targets from the fp32-rounded fp64 action
    a32 = out["d3_action_f64"].astype(np.float32)
    pub = oracle.clamp_mask(a32, d3_button0)
    qd, kp, kd = oracle.joint_targets(pub, d3_button0[:, None])
    out["d3_button0"] = d3_button0
    out["d3_published_from_f64"] = pub
    out["d3_qdes_from_f64"] = qd
    np.savez_compressed(os.path.join(HERE, "go2_policy_golden.npz"), **out)

    # closed loop (config 2 distribution), fp64 policy arithmetic
    n = 400
    raws = oracle.make_raw_states(n, seed=2)
    raws[3].joy_valid = 0
    raws[17].axes = np.array([0.0, 0.5, 0.1, -0.3], np.float32)       # axes[0]==0 -> -0.0f
    raws[29].quat = np.zeros(4, np.float32)                          # zero quaternion -> inverse() = 0
    st = oracle.ControllerState(H=2)
    cc = coracle.CController(cm, H=2)
    rec = {k: [] for k in ("obs", "action_raw", "action", "q_des", "kp")}
    for r in raws:
        so = oracle.controller_step(pol, st, r, np.float64)
        co = cc.step(coracle.raw_from_py(r), use_f64=True)
        assert np.array_equal(so.obs.view(np.uint32), np.frombuffer(co.obs, np.float32, 98).view(np.uint32)), "numpy vs C obs"
        assert np.array_equal(so.action.view(np.uint32), np.frombuffer(co.action, np.float32, 12).view(np.uint32))
        assert np.array_equal(so.q_des, np.frombuffer(co.q_des, np.float64, 12))
        for k in rec:
            rec[k].append(np.array(getattr(so, k)))
    raw_arr = {
        "quat": np.stack([r.quat for r in raws]), "gyro": np.stack([r.gyro for r in raws]),
        "q": np.stack([r.q for r in raws]), "dq": np.stack([r.dq for r in raws]),
        "axes": np.stack([r.axes for r in raws]), "foot_force": np.stack([r.foot_force for r in raws]),
        "joy_valid": np.array([r.joy_valid for r in raws], np.int32),
        "button0": np.array([r.button0 for r in raws], np.int32),
    }
    np.savez_compressed(os.path.join(HERE, "go2_closed_loop_golden.npz"),
                        **{f"raw_{k}": v for k, v in raw_arr.items()}, **{k: np.stack(v) for k, v in rec.items()})
    print("closed loop:", n, "steps; buttons pressed:", int(raw_arr["button0"].sum()))

    # the same raw-state stream through the reference's OWN controller.cpp / onnx_actor.cpp (oracle/_ref, compiled
    # from /root/reference against stub headers; A7 inside it is the C fp32 restatement) -> committed fixture, so the
    # restated pre/post-processing stays pinned to the reference's code where /root/reference does not exist
    if os.path.exists(coracle.REF_LIB_PATH) and os.path.exists(REF):
        rc = coracle.RefController()
        tr = {k: [] for k in ("obs", "action", "q_des", "kp", "kd")}
        for r in raws:
            axes = r.axes if r.joy_valid else np.zeros(0, np.float32)
            rc.feed(r.quat, r.gyro, r.foot_force, r.q.astype(np.float64), r.dq.astype(np.float64), axes, [int(r.button0)])
            out = rc.step()
            for k, v in zip(tr, out):
                tr[k].append(v.copy())
        rc.close()
        np.savez_compressed(os.path.join(HERE, "ref_controller_trace.npz"), **{k: np.stack(v) for k, v in tr.items()})
        print("reference-controller trace written:", len(raws), "steps")
    else:
        print("oracle/_ref not available: ref_controller_trace.npz left untouched")


if __name__ == "__main__":
    main()
