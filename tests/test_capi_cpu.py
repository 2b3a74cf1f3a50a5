"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/go2policy.h
declares, parses models on the host, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from go2_onnx_controller_b200 import actor, build, capi
from oracle import onnx_mini

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.fixture(scope="module")
def lib():
    build.build()
    return capi.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "go2policy.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(go2p_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (go2p_[a-z0-9_]+)", nm))
    assert declared <= exported


def test_abi_version_and_config_defaults(lib):
    assert lib.go2p_abi_version() == 2
    cfg = actor.default_config()
    # reference constants: controller.hpp:13-16,100-103,119-120,165; controller.cpp:244,246
    assert cfg.struct_size == C.sizeof(capi.Config)
    assert cfg.history == 2 and cfg.action_limit == 1000.0 and cfg.action_scale == 0.25
    assert list(cfg.q0) == [0.1, -0.1, 0.1, -0.1, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5]
    assert cfg.foot_threshold == 22 and cfg.kp == 28.0 and cfg.kd == 0.5 and cfg.kp_deadman == 5.0
    assert cfg.log_level == 2     # ORT_LOGGING_LEVEL_WARNING, onnx_actor.hpp:33


def test_struct_layouts_match_header():
    assert C.sizeof(capi.RawState) == 4 * 35 + 8 + 8
    assert C.sizeof(capi.StepOut) == 4 * (49 * 8) + 48 + 48 + 96 + 16 + 8
    assert C.sizeof(capi.MotorCmd) == 8 * 12 + 16


def test_motor_order_is_the_left_right_swap_of_the_foot_contacts(lib):
    """Unitree motor u = leg*3 + joint (legs FR, FL, RR, RL) <- Isaac joint*4 + leg (legs FL, FR, RL, RR,
    controller.hpp:168-170): the same [1,0,3,2] leg swap as the foot contacts (controller.hpp:100-103)."""
    order = (C.c_int32 * 12)()
    assert lib.go2p_motor_order(order) == capi.OK
    isaac = ["FL_hip", "FR_hip", "RL_hip", "RR_hip", "FL_thigh", "FR_thigh", "RL_thigh", "RR_thigh",
             "FL_calf", "FR_calf", "RL_calf", "RR_calf"]
    unitree = [f"{leg}_{j}" for leg in ("FR", "FL", "RR", "RL") for j in ("hip", "thigh", "calf")]
    assert [isaac[i] for i in order] == unitree
    assert sorted(order) == list(range(12))


def test_shard_rows_partition(lib):
    for B, n in ((1_048_576, 8), (10, 3), (0, 4), (5, 8)):
        prev = 0
        for i in range(n):
            b, e = actor.shard_rows(B, n, i)
            assert b == prev and e >= b and (e - b) in (B // n, B // n + 1)
            prev = e
        assert prev == B
    b, e = C.c_int64(), C.c_int64()
    assert lib.go2p_shard_rows(10, 0, 0, C.byref(b), C.byref(e)) == capi.ERR_INVALID


def _create(lib, path, cfg=None):
    h = C.c_void_p()
    cfg = cfg or actor.default_config()
    rc = lib.go2p_create(os.fspath(path).encode(), C.byref(cfg), C.byref(h))
    return rc, h, lib.go2p_last_error().decode()


def test_missing_model_is_io_error(lib, tmp_path):
    rc, h, msg = _create(lib, tmp_path / "nope.onnx")
    assert rc == capi.ERR_IO and not h and "nope.onnx" in msg


def test_garbage_model_is_model_error(lib, tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\x00\x01garbage that is not a protobuf" * 10)
    rc, h, msg = _create(lib, p)
    assert rc == capi.ERR_MODEL and not h and msg


def test_unsupported_op_is_rejected(lib, tmp_path):
    ws = [np.ones((4, 3), np.float32), np.ones((2, 4), np.float32)]
    bs = [np.zeros(4, np.float32), np.zeros(2, np.float32)]
    blob = onnx_mini.write_mlp_onnx(ws, bs, 1.0).replace(b"Elu", b"Erf")
    p = tmp_path / "erf.onnx"
    p.write_bytes(blob)
    rc, h, msg = _create(lib, p)
    assert rc == capi.ERR_MODEL and "Erf" in msg


def test_relu_and_observation_normaliser_are_folded_by_the_reader(lib, tmp_path):
    """Relu is served as Elu(alpha = 0); Sub / Div (and Add / Mul) with a constant in front of the first
    Gemm -- an observation normaliser -- are folded into that layer: W' = W * scale, b' = b + W @ shift."""
    rng = np.random.default_rng(4)
    dims = (20, 128, 128, 12)
    ws = [rng.normal(0, 1.0 / np.sqrt(k), (n, k)).astype(np.float32) for k, n in zip(dims[:-1], dims[1:])]
    bs = [rng.normal(0, 0.2, n).astype(np.float32) for n in dims[1:]]
    mean, std = rng.normal(0, 1, 20).astype(np.float32), rng.uniform(0.5, 2, 20).astype(np.float32)
    p = tmp_path / "norm_relu.onnx"
    p.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, 1.0, act_op="Relu", pre=[("Sub", mean), ("Div", std), ("Mul", np.float32([2.0]))]))
    info, cs = capi.ModelInfo(), C.c_double()
    assert lib.go2p_inspect_model(os.fspath(p).encode(), C.byref(info), C.byref(cs)) == capi.OK, lib.go2p_last_error()
    assert [info.dims[i] for i in range(4)] == list(dims)
    assert [info.has_elu[i] for i in range(3)] == [1, 1, 0] and info.elu_alpha[0] == 0.0 and info.elu_alpha[1] == 0.0
    scale = 2.0 / std.astype(np.float64)
    shift = -mean.astype(np.float64) / std.astype(np.float64) * 2.0
    w0 = (ws[0].astype(np.float64) * scale).astype(np.float32)
    b0 = (bs[0].astype(np.float64) + ws[0].astype(np.float64) @ shift).astype(np.float32)
    exp = 0.0
    for w, b in zip([w0] + ws[1:], [b0] + bs[1:]):
        flat = np.concatenate([w.reshape(-1), b]).astype(np.float64)
        exp += float(((np.arange(flat.size) % 97) + 1) @ flat)
    assert abs(cs.value - exp) <= 1e-9 * max(1.0, abs(exp))
    # the oracle evaluates the graph as written (normaliser nodes applied to the input, max(x, 0))
    from oracle import oracle
    pol = oracle.load_policy(str(p))
    x = rng.normal(0, 1, (7, 20)).astype(np.float32)
    h = (x.astype(np.float64) - mean) / std * 2.0
    for i, (w, b) in enumerate(zip(ws, bs)):
        h = h @ w.T.astype(np.float64) + b
        if i < 2:
            h = np.maximum(h, 0)
    assert np.abs(oracle.forward(pol, x) - h).max() <= 1e-12
    # a normaliser after the first layer, or constant / x, is not a normaliser: rejected with a message
    bad = onnx_mini.write_mlp_onnx(ws, bs, 1.0).replace(b"Elu", b"Div", 1)
    q = tmp_path / "bad.onnx"
    q.write_bytes(bad)
    assert lib.go2p_inspect_model(os.fspath(q).encode(), C.byref(info), None) == capi.ERR_MODEL


@pytest.mark.skipif(not _no_gpu(), reason="uses the no-device error to tell 'parsed' from 'rejected'")
@pytest.mark.parametrize("form,tail", [("matmul_add", False), ("mixed", True), ("gemm", True)])
def test_other_spellings_of_linear_layers_parse(lib, tmp_path, form, tail):
    """MatMul + Add, bias-less Gemm and a trailing Identity are accepted by the reader (the graph parses, so the
    create call gets as far as the device check); a stray Add is still rejected with a message."""
    rng = np.random.default_rng(1)
    ws = [rng.normal(0, 1, (8, 5)).astype(np.float32), rng.normal(0, 1, (3, 8)).astype(np.float32)]
    bs = [rng.normal(0, 1, 8).astype(np.float32), rng.normal(0, 1, 3).astype(np.float32)]
    p = tmp_path / "m.onnx"
    p.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, 1.0, form=form, identity_tail=tail))
    rc, h, msg = _create(lib, p)
    assert rc == capi.ERR_NO_DEVICE, msg
    pol = onnx_mini.load_policy(os.fspath(p))
    assert all(np.array_equal(l.weight, w) and np.array_equal(l.bias, b) for l, w, b in zip(pol.layers, ws, bs))
    # an Add that is not the bias of the preceding MatMul
    bad = onnx_mini.write_mlp_onnx(ws, bs, 1.0, form="matmul_add").replace(b"MatMul", b"Gemm\x00\x00")
    q = tmp_path / "bad.onnx"
    q.write_bytes(bad)
    rc, h, msg = _create(lib, q)
    assert rc == capi.ERR_MODEL and msg


def _checksum(policy):
    tot = 0.0
    for l in policy.layers:
        v = np.concatenate([l.weight.ravel(), l.bias.ravel()]).astype(np.float64)
        tot += float(((np.arange(v.size) % 97) + 1) @ v)
    return tot


def test_inspect_bundled_model_without_device(lib, model_path):
    """What print_model_info reports (reference: onnx_actor.cpp:60-66), parsed by the C++ reader with no GPU, and
    a positional checksum of every weight against the oracle-side reader of the same file."""
    info, cs = capi.ModelInfo(), C.c_double()
    assert lib.go2p_inspect_model(os.fspath(model_path).encode(), C.byref(info), C.byref(cs)) == capi.OK
    assert [info.dims[i] for i in range(info.n_layers + 1)] == [98, 128, 128, 128, 12]
    assert [info.has_elu[i] for i in range(4)] == [1, 1, 1, 0] and info.elu_alpha[0] == 1.0
    assert info.input_name == b"observation" and info.output_name == b"action" and info.n_params == 47244
    pol = onnx_mini.load_policy(os.fspath(model_path))
    assert abs(cs.value - _checksum(pol)) <= 1e-9 * max(1.0, abs(cs.value))


@pytest.mark.parametrize("kw", [
    dict(trans_b=False), dict(packed_dims=True), dict(use_float_data=True), dict(batch="batch"),
    dict(form="matmul_add"), dict(form="mixed", identity_tail=True), dict(final_activation=True),
])
def test_reader_agrees_with_oracle_reader_on_every_spelling(lib, tmp_path, kw):
    """Every serialisation variant the writer can produce parses to the same layers in the C++ reader (product) and
    the Python reader (oracle): dims, activations, alpha and the positional weight checksum."""
    rng = np.random.default_rng(len(str(kw)))
    dims = (23, 40, 17, 9)
    ws = [rng.normal(0, 1, (n, k)).astype(np.float32) for k, n in zip(dims[:-1], dims[1:])]
    bs = [rng.normal(0, 1, n).astype(np.float32) for n in dims[1:]]
    p = tmp_path / "m.onnx"
    p.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, 0.7, **kw))
    info, cs = capi.ModelInfo(), C.c_double()
    rc = lib.go2p_inspect_model(os.fspath(p).encode(), C.byref(info), C.byref(cs))
    assert rc == capi.OK, lib.go2p_last_error().decode()
    pol = onnx_mini.load_policy(os.fspath(p))
    assert [info.dims[i] for i in range(info.n_layers + 1)] == list(dims)
    assert [bool(info.has_elu[i]) for i in range(3)] == [l.elu_alpha is not None for l in pol.layers]
    assert all(abs(info.elu_alpha[i] - 0.7) < 1e-7 for i in range(3) if info.has_elu[i])
    assert abs(cs.value - _checksum(pol)) <= 1e-9 * max(1.0, abs(cs.value))
    assert all(np.array_equal(l.weight, w) and np.array_equal(l.bias, b) for l, w, b in zip(pol.layers, ws, bs))


def test_inspect_errors(lib, tmp_path):
    info = capi.ModelInfo()
    assert lib.go2p_inspect_model(os.fspath(tmp_path / "nope.onnx").encode(), C.byref(info), None) == capi.ERR_IO
    assert lib.go2p_inspect_model(None, C.byref(info), None) == capi.ERR_INVALID


def test_struct_size_mismatch_is_invalid(lib, model_path):
    cfg = actor.default_config()
    cfg.struct_size = 8
    rc, h, msg = _create(lib, model_path, cfg)
    assert rc == capi.ERR_INVALID and "struct_size" in msg


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device behaviour")
def test_no_device_means_no_compute(lib, model_path):
    """The product has no CPU fallback: a valid model without an sm_100 device fails loudly."""
    rc, h, msg = _create(lib, model_path)
    assert rc == capi.ERR_NO_DEVICE and not h
    assert "no CPU fallback" in msg or "sm_100a" in msg
    obs, act = np.zeros(98, np.float32), np.zeros(12, np.float32)
    with pytest.raises(capi.Go2PolicyError) as e:
        actor.ONNXActor(model_path, obs, act)
    assert e.value.code == capi.ERR_NO_DEVICE


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device behaviour")
def test_cpp_class_throws_without_device(model_path):
    """The reference-compatible C++ class (include/onnx_actor.hpp) surfaces the failure as std::exception,
    like the reference's Ort::Exception out of the constructor (onnx_actor.cpp:16)."""
    r = subprocess.run([build.SMOKE, model_path], capture_output=True, text=True)
    assert r.returncode == 1 and "ONNXActor" in r.stderr


def test_null_arguments_are_errors(lib):
    assert lib.go2p_act(None) == capi.ERR_INVALID
    assert lib.go2p_destroy(None) == capi.OK
    assert lib.go2p_infer_batch(None, None, None, 1, 0, None) == capi.ERR_INVALID
    assert lib.go2p_last_launch_count(None) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or open it."""
    pkg = os.path.join(ROOT, "go2_onnx_controller_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/" not in src, f


def test_bench_raw_state_generator_matches_both_struct_layouts():
    """bench.py generates its closed-loop inputs itself (the timed paths never touch oracle/); the C-ABI struct and the
    oracle's struct must be the same 156 bytes so the CPU leg can replay exactly the same raw states."""
    import bench
    from oracle import coracle
    arr = bench.synthetic_raw_states(capi, 16, seed=7)
    assert C.sizeof(capi.RawState) == C.sizeof(coracle.RawState) == 156
    for r in arr:
        o = coracle.RawState.from_buffer_copy(bytes(r))
        assert list(o.quat) == list(r.quat) and list(o.q) == list(r.q) and list(o.foot_force) == list(r.foot_force)
        assert o.joy_valid == r.joy_valid == 1 and o.button0 == r.button0
        assert abs(sum(v * v for v in r.quat) - 1.0) < 1e-5 and all(0 <= f <= 60 for f in r.foot_force)
    again = bench.synthetic_raw_states(capi, 16, seed=7)
    assert bytes(again) == bytes(arr)            # seeded: the same inputs on every rank and every run
