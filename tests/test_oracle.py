"""CPU tests of the oracle itself (test infrastructure): the numpy restatement, the C restatement and the
committed golden vectors must agree with each other and with the known-answer vectors of SURVEY.md App. D."""
import numpy as np
import pytest

from oracle import coracle, onnx_mini, oracle

KAT_ZEROS = [-0.795116175, 0.495336666, -0.280416397, 0.764204133, -0.693019657, -0.392156917,
             0.180990518, -0.362693726, 1.064260258, 0.517744812, 0.418805890, 0.866443300]
KAT_TWOS = [5.640298546, 0.859931734, 8.819978564, -3.643365861, -19.413944458, -7.316530415,
            0.820267543, 1.408699122, -1.644512823, -5.953271685, -2.744231396, -3.518555660]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_graph_structure(policy):
    # reference: onnx_inference/data/model.onnx (decoded)
    assert policy.dims == [98, 128, 128, 128, 12]
    assert [l.elu_alpha for l in policy.layers] == [1.0, 1.0, 1.0, None]
    assert policy.input_name == "observation" and policy.output_name == "action"
    assert policy.n_params == 47244


def test_known_answers(policy, golden):
    # smoke inputs of the reference: main.cpp:32 (zeros) and main.py:20 (2.0 * ones)
    np.testing.assert_allclose(oracle.forward(policy, np.zeros(98, np.float32)), KAT_ZEROS, rtol=0, atol=1e-9)
    np.testing.assert_allclose(oracle.forward(policy, 2 * np.ones(98, np.float32)), KAT_TWOS, rtol=0, atol=1e-8)
    np.testing.assert_allclose(golden["kat_action_f64"], [KAT_ZEROS, KAT_TWOS], rtol=0, atol=1e-8)


@pytest.mark.parametrize("name", ["d1", "d2"])
def test_numpy_oracle_matches_golden(policy, golden, name):
    y = oracle.forward(policy, golden[f"{name}_obs"], np.float64)
    np.testing.assert_allclose(y, golden[f"{name}_action_f64"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", ["d1", "d2"])
def test_c_oracle_matches_golden(cmodel, golden, name):
    X, ref = golden[f"{name}_obs"], golden[f"{name}_action_f64"]
    y64 = cmodel.forward_f64(X, threads=2)
    np.testing.assert_allclose(y64, ref, rtol=1e-12, atol=1e-12)
    for blocked in (False, True):
        y32 = cmodel.forward_f32(X, threads=2, blocked=blocked)
        err = np.abs(y32 - ref) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= 1e-5, err.max()     # the north-star fp32 budget; measured ~2e-6
        row = np.linalg.norm(y32 - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert row.max() <= 1e-5


def test_c_oracle_nonfinite_propagation(cmodel, golden):
    X, ref = golden["d3_obs"], golden["d3_action_f64"]
    with np.errstate(all="ignore"):
        y = cmodel.forward_f64(X)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(y[fin], ref[fin], rtol=1e-10, atol=1e-9)


def test_elu_semantics():
    # ONNX Elu-6 (Appendix C.8): x>=0 ? x : alpha*(exp(x)-1); -0.0 and NaN pass through
    x = np.array([-0.0, 0.0, -1.0, 2.0, np.nan, -np.inf, np.inf], np.float32)
    y = oracle.elu(x, 1.0)
    assert bits(y[0]) == 0x80000000 and y[1] == 0 and np.isnan(y[4]) and y[5] == -1.0 and y[6] == np.inf
    np.testing.assert_allclose(y[2], np.expm1(-1.0), rtol=1e-6)


def test_clamp_mask_semantics():
    # reference: controller.cpp:218-223; Appendix C.1-2
    a = np.array([np.nan, np.inf, -np.inf, -0.0, 1500.0, -1500.0, 3.5, -3.5], np.float32)
    c = oracle.clamp_mask(a, 0)
    assert np.isnan(c[0]) and c[1] == 1000 and c[2] == -1000 and bits(c[3]) == 0x80000000
    assert c[4] == 1000 and c[5] == -1000 and c[6] == np.float32(3.5)
    m = oracle.clamp_mask(a, 1)
    assert np.isnan(m[0])
    assert bits(m[1]) == 0 and bits(m[2]) == 0x80000000 and bits(m[7]) == 0x80000000 and bits(m[6]) == 0


def test_vel_cmd_and_joint_offsets():
    # Appendix C.3: axes[0]==0 -> -0.0f; double pow path
    v = oracle.vel_cmd_from_axes(np.array([0.0, 0.5, 9.0, -0.25], np.float32))
    assert bits(v[1]) == 0x80000000 and v[0] == 0.5 and v[2] == np.float32(-0.125)
    v = oracle.vel_cmd_from_axes(np.array([0.3, 0.5, 9.0, -0.25], np.float32))
    assert v[1] == np.float32(np.float64(np.float32(0.3)) ** 2 * 0.8)
    # Appendix C.4: q - q0 in double differs from the float path in the last bit for some inputs
    q = np.float32(0.1234567914)
    assert np.float32(np.float64(q) - 0.1) == np.float32(np.float64(q) - np.float64(0.1))


def test_gravity_identity_and_zero_quaternion():
    g = oracle.gravity_body(np.array([1, 0, 0, 0], np.float32))
    assert list(g) == [0.0, 0.0, -1.0]
    g = oracle.gravity_body(np.zeros(4, np.float32))      # Eigen inverse() of a zero quaternion is zero
    assert list(g) == [0.0, 0.0, -1.0]
    # 90 deg about x: body z axis points along world -y ... gravity in body frame = (0,-1,0)*sign
    s = np.float32(np.sqrt(0.5))
    g = oracle.gravity_body(np.array([s, s, 0, 0], np.float32))
    np.testing.assert_allclose(g, [0, -1, 0], atol=1e-6)


def test_observation_layout_term_major():
    # Appendix B: offsets 0,6,12,18,42,66,90; oldest frame first inside each block
    st = oracle.ControllerState(H=2)
    raws = oracle.make_raw_states(3, seed=11)
    obs = None
    cur = []
    for i, r in enumerate(raws):
        st.action = np.full(12, 10.0 + i, np.float32)
        obs = oracle.assemble_observation(st, r)
        cur.append(r)
    assert obs.shape == (98,)
    np.testing.assert_array_equal(obs[6:9], cur[1].gyro)
    np.testing.assert_array_equal(obs[9:12], cur[2].gyro)
    np.testing.assert_array_equal(obs[42:54], cur[1].dq)
    np.testing.assert_array_equal(obs[54:66], cur[2].dq)
    np.testing.assert_array_equal(obs[66:78], np.full(12, 11.0, np.float32))
    np.testing.assert_array_equal(obs[78:90], np.full(12, 12.0, np.float32))
    ff = cur[2].foot_force
    np.testing.assert_array_equal(obs[94:98], [float(ff[1] >= 22), float(ff[0] >= 22), float(ff[3] >= 22), float(ff[2] >= 22)])


def test_closed_loop_numpy_vs_c_vs_golden(policy, cmodel, golden_loop):
    g = golden_loop
    n = 120
    st = oracle.ControllerState(H=2)
    cc = coracle.CController(cmodel, H=2)
    for i in range(n):
        r = oracle.RawState(quat=g["raw_quat"][i], gyro=g["raw_gyro"][i], q=g["raw_q"][i], dq=g["raw_dq"][i],
                            foot_force=g["raw_foot_force"][i], axes=g["raw_axes"][i],
                            joy_valid=int(g["raw_joy_valid"][i]), button0=int(g["raw_button0"][i]))
        so = oracle.controller_step(policy, st, r, np.float64)
        co = cc.step(coracle.raw_from_py(r), use_f64=True)
        assert np.array_equal(bits(so.obs), bits(g["obs"][i])), i
        assert np.array_equal(bits(so.action), bits(g["action"][i])), i
        assert np.array_equal(so.q_des, g["q_des"][i]) and so.kp == g["kp"][i]
        assert np.array_equal(bits(np.frombuffer(co.obs, np.float32, 98)), bits(g["obs"][i])), i
        assert np.array_equal(bits(np.frombuffer(co.action, np.float32, 12)), bits(g["action"][i])), i
        assert np.array_equal(np.frombuffer(co.q_des, np.float64, 12), g["q_des"][i])
        assert co.kp == g["kp"][i] and co.kd == 0.5


def test_onnx_writer_reader_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    ws = [rng.standard_normal((7, 5)).astype(np.float32), rng.standard_normal((3, 7)).astype(np.float32)]
    bs = [rng.standard_normal(7).astype(np.float32), rng.standard_normal(3).astype(np.float32)]
    for kw in ({}, {"packed_dims": True}, {"use_float_data": True}, {"trans_b": False}, {"batch": "N"}):
        blob = onnx_mini.write_mlp_onnx(ws, bs, 0.5, **kw)
        pol = onnx_mini.load_policy(blob)
        assert pol.dims == [5, 7, 3]
        np.testing.assert_array_equal(pol.layers[0].weight, ws[0])
        np.testing.assert_array_equal(pol.layers[1].bias, bs[1])
        assert pol.layers[0].elu_alpha == 0.5 and pol.layers[1].elu_alpha is None
        p = tmp_path / "m.onnx"
        p.write_bytes(blob)
        cm = coracle.CModel(str(p))
        x = rng.standard_normal((4, 5)).astype(np.float32)
        np.testing.assert_allclose(cm.forward_f64(x), oracle.forward(pol, x), rtol=1e-12, atol=1e-12)


def test_operand_rounding_models_tensor_core_tolerances(policy, golden):
    # derives the stated tolerances of the 16-bit tensor-core path
    X, ref = golden["d2_obs"], golden["d2_action_f64"]
    e_bf16 = np.abs(oracle.forward_operand_rounded(policy, X, oracle.round_bf16) - ref).max()
    e_fp16 = np.abs(oracle.forward_operand_rounded(policy, X, oracle.round_fp16) - ref).max()
    assert e_bf16 < 5e-2 and e_fp16 < 1e-2, (e_bf16, e_fp16)
