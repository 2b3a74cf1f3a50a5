"""GPU parity tests (run on a B200 with -m gpu).  Every call goes through the C ABI of include/go2policy.h
(ctypes mirror in go2_onnx_controller_b200/capi.py); the oracle (oracle/) is only the checker.

Tolerances:
  fp32 path   |a-ref| <= 1e-5*max(|ref|,1) per element and row rel-L2 <= 1e-5      (ref = fp64 oracle)
  fp16 TC     <= 1e-2 abs on realistic observations (D2) and on N(0,1) (D1)
  bf16 TC     <= 3e-2 abs on D2, <= 7e-2 on D1
  clamp/mask/q_des/obs layout/history: bit-exact; gravity projection <= 1 ulp
"""
import ctypes as C
import re
import subprocess

import numpy as np
import pytest

from go2_onnx_controller_b200 import Go2Controller, ONNXActor, PolicyBatch, build, capi
from oracle import coracle, oracle

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

TOL_FP32 = 1e-5
# north_star: "a stated per-action tolerance, e.g. 1e-2 abs on normalised actions".  fp16 operands meet 1e-2 on both
# input sets (measured 1.9e-3 on D2, 6-8e-3 on D1); bf16 operands cannot (8-bit significand: 1.5e-2 / 4-5e-2 measured)
# and are offered with their own stated bound
TC_TOL = {capi.PREC_FP16: {"d2": 1e-2, "d1": 1e-2}, capi.PREC_BF16: {"d2": 3e-2, "d1": 7e-2}}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_fp32_parity(y, ref):
    fin = np.isfinite(ref)
    err = np.abs(y[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))
    assert err.max() <= TOL_FP32, f"fp32 parity {err.max():.3e}"
    if ref.ndim == 2 and fin.all():
        row = np.linalg.norm(y - ref, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-30)
        assert row.max() <= TOL_FP32, f"row rel-L2 {row.max():.3e}"


def raw_struct(g, i):
    r = capi.RawState()
    r.quat[:] = g["raw_quat"][i].tolist(); r.gyro[:] = g["raw_gyro"][i].tolist()
    r.q[:] = g["raw_q"][i].tolist(); r.dq[:] = g["raw_dq"][i].tolist(); r.axes[:] = g["raw_axes"][i].tolist()
    r.foot_force[:] = [int(v) for v in g["raw_foot_force"][i]]
    r.joy_valid = int(g["raw_joy_valid"][i]); r.button0 = int(g["raw_button0"][i])
    return r


def raw_py(g, i):
    return oracle.RawState(quat=g["raw_quat"][i], gyro=g["raw_gyro"][i], q=g["raw_q"][i], dq=g["raw_dq"][i],
                           foot_force=g["raw_foot_force"][i], axes=g["raw_axes"][i],
                           joy_valid=int(g["raw_joy_valid"][i]), button0=int(g["raw_button0"][i]))


# ------------------------------------------------------------------------------------------ batch 1
@pytest.mark.parametrize("mode", [capi.B1_PERSISTENT, capi.B1_GRAPH, capi.B1_LAUNCH])
def test_act_known_answers_all_modes(torch_cuda, model_path, golden, mode):
    """ONNXActor::act() semantics (reference: onnx_actor.cpp:38-48): reads the bound observation buffer's current
    contents, overwrites the bound action buffer; zeros and 2*ones are the reference's smoke inputs."""
    obs, act = np.zeros(98, np.float32), np.zeros(12, np.float32)
    a = ONNXActor(model_path, obs, act, b1_mode=mode)
    try:
        assert a.check_dims()
        for x, ref in zip(golden["kat_obs"], golden["kat_action_f64"]):
            obs[:] = x
            act[:] = 7.0
            a.act()
            assert_fp32_parity(act.astype(np.float64), ref)
        for x, ref in zip(golden["d1_obs"][:64], golden["d1_action_f64"][:64]):
            obs[:] = x
            a.act()
            assert_fp32_parity(act.astype(np.float64), ref)
        st = a.stats()
        assert st.steps == 66 and 0 < st.device_ns_min < 1_000_000
    finally:
        a.close()


def test_act_nonfinite_and_signed_zero(torch_cuda, model_path, policy):
    obs, act = np.zeros(98, np.float32), np.zeros(12, np.float32)
    a = ONNXActor(model_path, obs, act)
    try:
        obs[:] = np.nan
        a.act()
        assert np.isnan(act).all()
        obs[:] = -0.0
        a.act()
        assert_fp32_parity(act.astype(np.float64), oracle.forward(policy, np.zeros(98, np.float32)))
    finally:
        a.close()


def test_cpp_class_smoke_binary(torch_cuda, model_path, golden):
    """include/onnx_actor.hpp through the equivalent of the reference's smoke main (main.cpp:24-48)."""
    r = subprocess.run([build.SMOKE, model_path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    # print_model_info: the reference's four lines (onnx_actor.cpp:60-66)
    assert "Input dimension: 98" in out and "Output dimension: 12" in out
    assert "Input name: observation" in out and "Output name: action" in out
    vals = [float(v) for v in re.search(r"Action: \[(.*)\]", out).group(1).split(",")]
    np.testing.assert_allclose(vals, golden["kat_action_f64"][0], atol=2e-5)


def test_fused_step_closed_loop_vs_oracle(torch_cuda, model_path, policy, golden_loop):
    """A1-A6 + A7 + A9 + A11 in one resident-kernel round trip vs the restated publish()
    (reference: controller.cpp:173-251).  The oracle is re-synchronised to the device's own raw action each step
    (so every step is an exact comparison) and the free-running golden trajectory bounds the drift."""
    g = golden_loop
    n = g["obs"].shape[0]
    # trace recorded from the reference's own compiled controller.cpp (tests/golden/make_golden.py, oracle/_ref)
    ref_tr = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_controller_trace.npz"))
    ctl = Go2Controller(model_path)
    st = oracle.ControllerState(H=2)
    try:
        max_err, max_drift, ulp_g = 0.0, 0.0, 0
        for i in range(n):
            out = ctl.step(raw_struct(g, i))
            d_obs = np.frombuffer(out.observation, np.float32, 98).copy()
            d_raw = np.frombuffer(out.action_raw, np.float32, 12).copy()
            d_act = np.frombuffer(out.action, np.float32, 12).copy()
            d_qdes = np.frombuffer(out.q_des, np.float64, 12).copy()
            so = oracle.controller_step(policy, st, raw_py(g, i), np.float64, act_fn=lambda o: d_raw)
            # gravity (obs[0:6]) <= 1 ulp (Appendix C.6), everything else bit-exact
            ulp = np.abs(bits(d_obs[:6]).astype(np.int64) - bits(so.obs[:6]).astype(np.int64)).max()
            ulp_g = max(ulp_g, int(ulp))
            assert ulp <= 1, (i, d_obs[:6], so.obs[:6])
            assert np.array_equal(bits(d_obs[6:]), bits(so.obs[6:])), i
            # everything in the observation that does not pass through the policy (all but the previous-action
            # block 66..89) must equal what the reference's compiled publish() produced, bit for bit
            ro = ref_tr["obs"][i]
            assert np.array_equal(bits(d_obs[6:66]), bits(ro[6:66])) and np.array_equal(bits(d_obs[90:]), bits(ro[90:])), i
            assert np.abs(bits(d_obs[:6]).astype(np.int64) - bits(ro[:6]).astype(np.int64)).max() <= 1, i
            assert np.abs(d_obs[66:90] - ro[66:90]).max() <= 1e-4, i
            assert np.array_equal(bits(d_act), bits(so.action)), i                    # A9 bit-exact
            assert np.array_equal(d_qdes, so.q_des), i                                # A11 bit-exact (double)
            assert out.kp == so.kp and out.kd == so.kd, i
            ref = oracle.forward(policy, d_obs, np.float64)
            max_err = max(max_err, float((np.abs(d_raw - ref) / np.maximum(1, np.abs(ref))).max()))
            max_drift = max(max_drift, float(np.abs(d_act - g["action"][i]).max()))
            assert 0 < out.device_ns < 1_000_000
            st.hist[0][:] = d_obs[:6]      # keep the <=1-ulp gravity difference from accumulating in the checker
        assert max_err <= TOL_FP32, max_err
        assert max_drift <= 1e-3, max_drift
        print(f"closed loop: {n} steps, max fp32 err {max_err:.2e}, drift vs fp64 golden {max_drift:.2e}, gravity ulp {ulp_g}")
    finally:
        ctl.close()


def test_fused_step_reset_gains_and_modes(torch_cuda, model_path, golden_loop):
    g = golden_loop
    outs = {}
    for mode in (capi.B1_PERSISTENT, capi.B1_GRAPH, capi.B1_LAUNCH):
        ctl = Go2Controller(model_path, b1_mode=mode)
        try:
            seq = []
            for rep in range(2):
                for i in range(20):
                    o = ctl.step(raw_struct(g, i))
                    seq.append(np.frombuffer(o.action, np.float32, 12).copy())
                ctl.reset()                               # controller.hpp:132-162 initial state
            for i, (a, b) in enumerate(zip(seq[:20], seq[20:])):
                assert np.array_equal(bits(a), bits(b)), f"mode {mode}: step {i} differs after reset: {a} vs {b}"
            ctl.set_gains(31.5, 0.75)                     # controller.cpp:254-277
            r = raw_struct(g, 0)
            o = ctl.step(r)
            assert o.kp == 31.5 and o.kd == 0.75
            r.button0 = 1
            o = ctl.step(r)
            assert o.kp == 5.0 and o.kd == 0.75           # controller.cpp:246
            assert all(v == 0.0 for v in o.action)
            outs[mode] = np.stack(seq)
        finally:
            ctl.close()
    assert np.array_equal(bits(outs[capi.B1_PERSISTENT]), bits(outs[capi.B1_GRAPH]))
    assert np.array_equal(bits(outs[capi.B1_PERSISTENT]), bits(outs[capi.B1_LAUNCH]))


def test_selfdriven_profiling_twin_matches_resident_kernel(torch_cuda, model_path, golden_loop):
    """The bounded launch ncu profiles runs the same device functions as the resident kernel: after n closed-loop
    steps both hold the same published action, bit for bit."""
    g = golden_loop
    n = 64
    raws = [raw_struct(g, i) for i in range(n)]
    ctl = Go2Controller(model_path)
    try:
        _, _, last = ctl.closed_loop(raws, n)
        ctl.stop()
        act, ms = ctl.selfdriven(raws, n)
        assert np.array_equal(bits(act), bits(np.frombuffer(last.action, np.float32, 12)))
        assert 0 < ms < 100
    finally:
        ctl.close()


def test_resident_kernel_idle_farewell_and_relaunch(torch_cuda, model_path, golden):
    """The resident kernel leaves its SM after idle_exit_ms and is relaunched transparently."""
    import time
    ctl = Go2Controller(model_path, idle_exit_ms=200)
    try:
        g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "go2_closed_loop_golden.npz"))
        a0 = np.frombuffer(ctl.step(raw_struct(g, 0)).action, np.float32, 12).copy()
        time.sleep(0.6)
        a1 = np.frombuffer(ctl.step(raw_struct(g, 1)).action, np.float32, 12).copy()
        assert np.abs(a0 - g["action"][0]).max() < 1e-3 and np.abs(a1 - g["action"][1]).max() < 1e-3
    finally:
        ctl.close()


# ------------------------------------------------------------------------------------------ batched
@pytest.fixture(scope="module")
def pb(torch_cuda, model_path):
    p = PolicyBatch(model_path)
    yield p
    p.close()


def run_batch(torch, pb, X, prec, button0=None, flags=0):
    B = X.shape[0]
    d_obs = torch.from_numpy(np.ascontiguousarray(X, np.float32)).cuda()
    d_act = torch.full((max(B, 1), pb.out_dim), 777.0, device="cuda", dtype=torch.float32)
    d_q = torch.zeros((max(B, 1), 12), device="cuda", dtype=torch.float64) if flags & capi.F_QDES else None
    d_b = torch.from_numpy(np.ascontiguousarray(button0, np.int32)).cuda() if button0 is not None else None
    torch.cuda.synchronize()
    pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), B, prec, 0, d_b.data_ptr() if d_b is not None else None,
                    d_q.data_ptr() if d_q is not None else None, flags)
    torch.cuda.synchronize()
    return d_act[:B].cpu().numpy(), (d_q[:B].cpu().numpy() if d_q is not None else None)


@pytest.mark.parametrize("name", ["kat", "d1", "d2"])
def test_batched_fp32_vs_golden(torch_cuda, pb, golden, name):
    y, _ = run_batch(torch_cuda, pb, golden[f"{name}_obs"], capi.PREC_FP32)
    assert_fp32_parity(y.astype(np.float64), golden[f"{name}_action_f64"])
    assert pb.last_launches() == 4


@pytest.mark.parametrize("prec", [capi.PREC_FP16, capi.PREC_BF16])
@pytest.mark.parametrize("name", ["d1", "d2"])
def test_batched_tensor_core_vs_golden(torch_cuda, pb, golden, prec, name):
    assert pb.info.tensor_core_path == 1
    y, _ = run_batch(torch_cuda, pb, golden[f"{name}_obs"], prec)
    err = np.abs(y - golden[f"{name}_action_f64"]).max()
    print(f"TC prec={prec} {name}: max abs err {err:.3e}")
    assert err <= TC_TOL[prec][name], err
    assert pb.last_launches() == 1          # the whole chain is ONE kernel


@pytest.mark.parametrize("prec", [capi.PREC_FP32, capi.PREC_FP16, capi.PREC_BF16])
@pytest.mark.parametrize("B", [1, 2, 127, 128, 129, 255, 1000, 4096, 148 * 128 * 2 + 77])
def test_batched_ragged_sizes(torch_cuda, pb, cmodel, prec, B):
    X = oracle.make_obs_d1(B, 98, seed=B)
    ref = cmodel.forward_f64(X, threads=8)
    y, _ = run_batch(torch_cuda, pb, X, prec)
    if prec == capi.PREC_FP32:
        assert_fp32_parity(y.astype(np.float64), ref)
    else:
        # the maximum over up to 38k x 12 actions of N(0,1) inputs is a tail statistic: 1.5x the 512-row golden bound
        assert np.abs(y - ref).max() <= 1.5 * TC_TOL[prec]["d1"]


def test_batched_empty_is_noop(torch_cuda, pb):
    for prec in (capi.PREC_FP32, capi.PREC_FP16):
        y, _ = run_batch(torch_cuda, pb, np.zeros((0, 98), np.float32), prec)
        assert y.shape == (0, 12) and pb.last_launches() == 0


@pytest.mark.parametrize("prec", [capi.PREC_FP32, capi.PREC_FP16, capi.PREC_BF16])
def test_clamp_mask_qdes_bit_exact_on_stress_set(torch_cuda, pb, golden, prec):
    """A9/A11 epilogue (reference: controller.cpp:217-223,244): NaN passes the clamp, +-Inf -> +-1000, masked
    negatives become -0.0; q_des = q0 + (double)a*0.25 in double.  Bit-exact given the kernel's own raw action."""
    X, b0 = golden["d3_obs"], golden["d3_button0"]
    raw, _ = run_batch(torch_cuda, pb, X, prec)
    pub, qd = run_batch(torch_cuda, pb, X, prec, button0=b0, flags=capi.F_CLAMP_MASK | capi.F_QDES)
    ref_raw = golden["d3_action_f64"]
    if prec == capi.PREC_FP16:
        # fp16 operands saturate at +-65504 (cvt.satfinite): +-Inf observations and the 1e5-sized activations of
        # this x3000 set are outside the format, so only NaN propagation is comparable (NaN stays NaN)
        nan_in = np.isnan(X).any(axis=1)
        assert np.isnan(raw[nan_in]).all()
    else:
        assert np.array_equal(np.isnan(raw), np.isnan(ref_raw))
        assert (np.abs(raw) > 1000).mean() > 0.1           # the clip really fires on this set
    exp_pub = oracle.clamp_mask(raw, b0)
    assert np.array_equal(bits(pub), bits(exp_pub))
    exp_qd, _, _ = oracle.joint_targets(exp_pub, b0[:, None])
    assert np.array_equal(qd.view(np.uint64), exp_qd.view(np.uint64))
    if prec == capi.PREC_FP32:
        fin = np.isfinite(ref_raw)
        # x3000 inputs: hidden activations reach 1e6 and cancel down to 1e3-sized actions, so fp32 accumulation
        # error is relative to the activations, not the action (the C fp32 oracle shows the same 1e-4..1e-3)
        assert (np.abs(raw[fin] - ref_raw[fin]) / np.maximum(1, np.abs(ref_raw[fin]))).max() <= 2e-3
        # and against the committed fixture: wherever the fp64 action is clipped or masked (and not within
        # rounding of the clip edge) the published bits must be the fixture's +-1000 / +-0.0 exactly
        gold = golden["d3_published_from_f64"]
        decided = fin & (np.abs(np.abs(ref_raw) - 1000) > 5.0) & ((np.abs(gold) == 1000) | (gold == 0))
        assert decided.sum() > 500
        assert np.array_equal(bits(pub[decided]), bits(gold[decided]))


def test_batched_null_button_means_unmasked(torch_cuda, pb, golden):
    X = golden["d3_obs"]
    raw, _ = run_batch(torch_cuda, pb, X, capi.PREC_FP16)
    pub, _ = run_batch(torch_cuda, pb, X, capi.PREC_FP16, flags=capi.F_CLAMP_MASK)
    assert np.array_equal(bits(pub), bits(oracle.clamp_mask(raw, 0)))


def test_host_buffer_pipeline_matches_device_path(torch_cuda, pb, cmodel):
    B = 200_001
    X = oracle.make_obs_d1(B, 98, seed=9)
    for prec in (capi.PREC_FP16, capi.PREC_FP32):
        y_dev, _ = run_batch(torch_cuda, pb, X, prec)
        hx = pb.pinned((B, 98)); hy = pb.pinned((B, 12))
        hx[:] = X
        y_host = pb.infer_host(hx, hy, prec)
        assert np.array_equal(bits(y_host), bits(y_dev))
        y_pageable = pb.infer_host(X, None, prec)
        assert np.array_equal(bits(y_pageable), bits(y_dev))
    assert_fp32_parity(y_dev[:4096].astype(np.float64), cmodel.forward_f64(X[:4096], 8))


@pytest.mark.parametrize("prec", [capi.PREC_FP16, capi.PREC_BF16, capi.PREC_FP32])
def test_full_size_rows_are_batch_independent(torch_cuda, pb, cmodel, prec):
    """BASELINE.json configs[3] size (1,048,576 rows): a row's action does not depend on the batch it travels in
    -- random rows recomputed alone (small batch) are bit-identical, and a sample agrees with the oracle."""
    torch = torch_cuda
    B = 1_048_576
    g = torch.Generator(device="cuda").manual_seed(0)
    d_obs = torch.randn((B, 98), device="cuda", dtype=torch.float32, generator=g)
    d_act = torch.empty((B, 12), device="cuda", dtype=torch.float32)
    pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), B, prec)
    torch.cuda.synchronize()
    idx = torch.randint(0, B, (3000,), device="cuda", generator=g)
    sub = d_obs[idx].contiguous()
    sub_act = torch.empty((3000, 12), device="cuda", dtype=torch.float32)
    pb.infer_device(sub.data_ptr(), sub_act.data_ptr(), 3000, prec)
    torch.cuda.synchronize()
    assert torch.equal(d_act[idx].view(torch.int32), sub_act.view(torch.int32))
    ref = cmodel.forward_f64(sub.cpu().numpy(), threads=8)
    err = np.abs(sub_act.cpu().numpy() - ref)
    if prec == capi.PREC_FP32:
        assert (err / np.maximum(1, np.abs(ref))).max() <= TOL_FP32
    else:
        assert err.max() <= TC_TOL[prec]["d1"]
    # checksum of checksums: a second full pass is bit-identical (no race, no stale tile)
    d_act2 = torch.empty_like(d_act)
    pb.infer_device(d_obs.data_ptr(), d_act2.data_ptr(), B, prec)
    torch.cuda.synchronize()
    assert torch.equal(d_act.view(torch.int32), d_act2.view(torch.int32))


def test_batched_observation_assembly_bit_exact(torch_cuda, pb, golden_loop):
    """A1-A6 for B robots on the device vs the oracle, 4 consecutive steps with per-robot history."""
    torch = torch_cuda
    g = golden_loop
    B = 96
    steps = 4
    states = [oracle.ControllerState(H=2) for _ in range(B)]
    d_obs = torch.zeros((B, 98), device="cuda", dtype=torch.float32)
    d_vel = torch.zeros((B, 3), device="cuda", dtype=torch.float32)
    rng = np.random.default_rng(4)
    for s in range(steps):
        raws = (capi.RawState * B)()
        prev = rng.normal(0, 1.5, (B, 12)).astype(np.float32)
        exp = np.zeros((B, 98), np.float32)
        for b in range(B):
            i = (s * B + b) % g["obs"].shape[0]
            raws[b] = raw_struct(g, i)
            states[b].action = prev[b].copy()
            exp[b] = oracle.assemble_observation(states[b], raw_py(g, i))
        d_raw = torch.from_numpy(np.frombuffer(bytes(raws), np.uint8).copy()).cuda()
        d_prev = torch.from_numpy(prev).cuda()
        pb.assemble_device(d_raw.data_ptr(), d_prev.data_ptr(), d_vel.data_ptr(), d_obs.data_ptr(), B)
        torch.cuda.synchronize()
        got = d_obs.cpu().numpy()
        for b in range(B):
            ulp = np.abs(bits(got[b, :6]).astype(np.int64) - bits(exp[b, :6]).astype(np.int64)).max()
            assert ulp <= 1
            states[b].hist[0][:] = got[b, :6]
        assert np.array_equal(bits(got[:, 6:]), bits(exp[:, 6:])), s


@pytest.mark.parametrize("prec", [capi.PREC_FP32, capi.PREC_FP16])
def test_batched_controller_step_closed_loop(torch_cuda, pb, policy, golden_loop, prec):
    """go2p_step_batch = publish() for B robots (one fused launch on the tensor-core path; assembly + policy launches
    on the fp32 path), per-robot history on the device, the published action fed back as the next step's previous action (reference: controller.cpp:173-251).
    Every step is compared with the restated publish() re-synchronised to the device's own published action."""
    torch = torch_cuda
    g = golden_loop
    B, steps = 333, 6           # two full 128-row tiles + a ragged one
    n_g = g["obs"].shape[0]
    states = [oracle.ControllerState(H=2) for _ in range(B)]
    d_obs = torch.zeros((B, 98), device="cuda", dtype=torch.float32)
    d_vel = torch.zeros((B, 3), device="cuda", dtype=torch.float32)
    d_act = torch.zeros((B, 12), device="cuda", dtype=torch.float32)
    d_q = torch.zeros((B, 12), device="cuda", dtype=torch.float64)
    tol = TOL_FP32 if prec == capi.PREC_FP32 else 1e-2
    masked = 0
    for s in range(steps):
        raws = (capi.RawState * B)()
        idx = [(7 * s * B + 3 * b) % n_g for b in range(B)]
        for b in range(B):
            raws[b] = raw_struct(g, idx[b])
            if (b + s) % 9 == 0:
                raws[b].button0 = 1          # dead-man pressed on some robots
        d_raw = torch.from_numpy(np.frombuffer(bytes(raws), np.uint8).copy()).cuda()
        pb.step_device(d_raw.data_ptr(), d_vel.data_ptr(), d_obs.data_ptr(), d_act.data_ptr(), d_q.data_ptr(), B, prec)
        torch.cuda.synchronize()
        # fp32: assembly + 3 Gemm launches + output kernel; tensor-core precisions: ONE launch (A1-A6 fused into the
        # policy kernel's conversion job)
        assert pb.last_launches() == (5 if prec == capi.PREC_FP32 else 1)
        obs, act, qd = d_obs.cpu().numpy(), d_act.cpu().numpy(), d_q.cpu().numpy()
        for b in range(B):
            raw = raw_py(g, idx[b])
            raw.button0 = int(raws[b].button0)
            so = oracle.controller_step(policy, states[b], raw, np.float64, act_fn=lambda o: oracle.forward(policy, obs[b], np.float64))
            assert np.abs(bits(obs[b, :6]).astype(np.int64) - bits(so.obs[:6]).astype(np.int64)).max() <= 1, (s, b)
            assert np.array_equal(bits(obs[b, 6:]), bits(so.obs[6:])), (s, b)
            assert (np.abs(act[b] - so.action) / np.maximum(1, np.abs(so.action))).max() <= tol, (s, b)
            if raw.button0:
                masked += 1
                assert not act[b].any()                                   # masked to (signed) zero
            exp_q, _, _ = oracle.joint_targets(act[b], raw.button0)
            assert np.array_equal(qd[b].view(np.uint64), exp_q.view(np.uint64)), (s, b)   # A11 bit-exact (double)
            states[b].action = act[b].copy()           # closed loop on the device's own published action
            states[b].hist[0][:] = obs[b, :6]          # keep the <= 1-ulp gravity difference from accumulating
    assert masked > 20


def test_resident_kernel_coexists_with_batched_kernels(torch_cuda, model_path, golden):
    """One handle: the batch-1 resident kernel keeps answering while batched launches share the device."""
    obs, act = np.zeros(98, np.float32), np.zeros(12, np.float32)
    a = ONNXActor(model_path, obs, act)
    p = PolicyBatch(model_path)
    try:
        a.act()
        X = golden["d1_obs"]
        for _ in range(3):
            y, _ = run_batch(torch_cuda, p, X, capi.PREC_FP16)
            obs[:] = X[0]
            a.act()
            assert_fp32_parity(act.astype(np.float64), golden["d1_action_f64"][0])
            assert np.abs(y - golden["d1_action_f64"]).max() <= TC_TOL[capi.PREC_FP16]["d1"]
    finally:
        a.close()
        p.close()


@pytest.mark.parametrize("shape", [
    dict(dims=(77, 128, 128, 7), alpha=0.5, final_act=False),        # odd input width, 3 layers, 7 outputs
    dict(dims=(64, 128, 12), alpha=1.0, final_act=True),             # 2 layers, ELU on the output layer too
    dict(dims=(40, 128, 128, 128, 128, 16), alpha=1.3, final_act=False),    # 5 layers, widest output the TC kernel serves
    dict(dims=(126, 128, 128, 12), alpha=1.0, final_act=False),      # widest input of the one-kernel path (126 + 2 ones = 128)
    dict(dims=(140, 128, 128, 12), alpha=1.0, final_act=False),      # wider input: served by the per-layer GEMM kernels
])
def test_other_narrow_policies_all_paths(torch_cuda, tmp_path, shape):
    """The kernels are driven by the parsed graph, not by the bundled policy's constants: other layer counts, input
    widths (odd: scalar conversion path), output widths (generic output epilogue), ELU alphas and an activation on the
    last layer (scaled-domain output) must agree with the oracle on every path that serves them."""
    from oracle import onnx_mini
    dims = shape["dims"]
    rng = np.random.default_rng(sum(dims))
    ws = [rng.normal(0, 1.0 / np.sqrt(k), (n, k)).astype(np.float32) for k, n in zip(dims[:-1], dims[1:])]
    bs = [rng.normal(0, 0.2, n).astype(np.float32) for n in dims[1:]]
    path = tmp_path / "m.onnx"
    path.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, shape["alpha"], final_activation=shape["final_act"]))
    cm = coracle.CModel(str(path))
    p = PolicyBatch(str(path))
    try:
        assert p.info.tensor_core_path == 1
        for B in (1, 129, 1000):
            X = rng.normal(0, 1, (B, dims[0])).astype(np.float32)
            ref = cm.forward_f64(X, 4)
            y, _ = run_batch(torch_cuda, p, X, capi.PREC_FP32)
            assert_fp32_parity(y.astype(np.float64), ref)
            y, _ = run_batch(torch_cuda, p, X, capi.PREC_FP16)
            assert np.abs(y - ref).max() <= 3e-2, np.abs(y - ref).max()
            y, _ = run_batch(torch_cuda, p, X, capi.PREC_BF16)
            assert np.abs(y - ref).max() <= 1.5e-1, np.abs(y - ref).max()
        # batch-1 act() on the same graph (generic kernel: weights in shared memory)
        obs, act = np.zeros(dims[0], np.float32), np.zeros(dims[-1], np.float32)
        a = ONNXActor(str(path), obs, act)
        try:
            for i in range(5):
                obs[:] = X[i]
                a.act()
                assert_fp32_parity(act.astype(np.float64), ref[i])
        finally:
            a.close()
    finally:
        p.close()


def test_matmul_add_spelling_gives_identical_bits(torch_cuda, tmp_path, policy):
    """The same weights written as Gemm or as MatMul + Add (+ trailing Identity) are the same policy to the kernels."""
    from oracle import onnx_mini
    ws = [l.weight for l in policy.layers]
    bs = [l.bias for l in policy.layers]
    X = oracle.make_obs_d1(777, 98, seed=12)
    outs = []
    for form, tail in (("gemm", False), ("matmul_add", True), ("mixed", False)):
        path = tmp_path / f"{form}.onnx"
        path.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, policy.layers[0].elu_alpha, form=form, identity_tail=tail))
        p = PolicyBatch(str(path))
        try:
            outs.append([run_batch(torch_cuda, p, X, prec)[0] for prec in (capi.PREC_FP32, capi.PREC_FP16)])
        finally:
            p.close()
    for o in outs[1:]:
        assert np.array_equal(bits(o[0]), bits(outs[0][0])) and np.array_equal(bits(o[1]), bits(outs[0][1]))
    assert_fp32_parity(outs[0][0].astype(np.float64), oracle.forward(policy, X, np.float64))


def test_wide_policy_fp32_path(torch_cuda, wide_model_path):
    """BASELINE.json configs[4]: synthetic 245-1024-512-256-12 ELU policy (5-frame history input)."""
    cm = coracle.CModel(wide_model_path)
    p = PolicyBatch(wide_model_path, history=5)
    try:
        assert [p.info.dims[i] for i in range(5)] == [245, 1024, 512, 256, 12]
        X = oracle.make_obs_d1(3001, 245, seed=5)
        y, _ = run_batch(torch_cuda, p, X, capi.PREC_FP32)
        assert_fp32_parity(y.astype(np.float64), cm.forward_f64(X, 8))
    finally:
        p.close()


@pytest.mark.parametrize("prec", [capi.PREC_FP16, capi.PREC_BF16])
def test_wide_policy_tensor_core_path(torch_cuda, wide_model_path, prec):
    """configs[4] on the per-layer tcgen05 GEMM kernels (kernels_wide.cuh): ragged batches, a batch larger than one
    pass (4 x 148 row tiles = 75,776 rows), the fused A9/A11 epilogue, and the host-buffer pipeline (concurrent streams)."""
    cm = coracle.CModel(wide_model_path)
    p = PolicyBatch(wide_model_path, history=5)
    tol = 2e-2 if prec == capi.PREC_FP16 else 1.2e-1
    try:
        assert p.info.tensor_core_path == 1
        CH = 4 * 148 * 128
        for B in (1, 129, 3001, CH + 77):
            X = oracle.make_obs_d1(B, 245, seed=B)
            ref = cm.forward_f64(X, 8)
            y, _ = run_batch(torch_cuda, p, X, prec)
            err = np.abs(y - ref).max()
            print(f"wide TC prec={prec} B={B}: max abs err {err:.3e} (|ref| max {np.abs(ref).max():.2f})")
            assert err <= tol, err
            assert p.last_launches() == 5 * ((B + CH - 1) // CH)
        # rows do not depend on the batch they are in (same tile arithmetic everywhere)
        y1, _ = run_batch(torch_cuda, p, X[CH - 100:CH + 200], prec)
        assert np.array_equal(bits(y1), bits(y[CH - 100:CH + 200]))
        # A9 / A11 epilogue bit-exact given the kernel's own raw action
        b0 = (np.arange(B) % 3 == 0).astype(np.int32)
        pub, qd = run_batch(torch_cuda, p, X * 50.0, prec, button0=b0, flags=capi.F_CLAMP_MASK | capi.F_QDES)
        raw, _ = run_batch(torch_cuda, p, X * 50.0, prec)
        exp_pub = oracle.clamp_mask(raw, b0)
        assert np.array_equal(bits(pub), bits(exp_pub))
        exp_qd, _, _ = oracle.joint_targets(exp_pub, b0[:, None])
        assert np.array_equal(qd.view(np.uint64), exp_qd.view(np.uint64))
        # host buffers: three streams in flight, each with its own activation buffers
        Bh = 150_000
        Xh = oracle.make_obs_d1(Bh, 245, seed=3)
        y_dev, _ = run_batch(torch_cuda, p, Xh, prec)
        y_host = p.infer_host(Xh, None, prec)
        assert np.array_equal(bits(y_host), bits(y_dev))
    finally:
        p.close()


def test_unknown_precision_fails_loudly(torch_cuda, pb, golden):
    """3 was reserved for kind::tf32, which is not built (include/go2policy.h): it is rejected, never approximated."""
    for bad in (3, 7, -1):
        with pytest.raises(capi.Go2PolicyError) as e:
            run_batch(torch_cuda, pb, golden["kat_obs"], bad)
        assert e.value.code == capi.ERR_INVALID
