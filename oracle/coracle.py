"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_build/liboracle.so (the plain-C
restatement) and of oracle/_ref/libref_controller.so (the reference's own controller.cpp /
onnx_actor.cpp compiled against the stub headers in oracle/ref_stubs/, see oracle/Makefile).
Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libref_controller.so")

ORC_MAX_LAYERS = 16
ORC_MAX_HIST = 8
ORC_FRAME = 49


class OrcModel(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int),
        ("dims", C.c_int * (ORC_MAX_LAYERS + 1)),
        ("w", C.POINTER(C.c_float) * ORC_MAX_LAYERS),
        ("wt", C.POINTER(C.c_float) * ORC_MAX_LAYERS),
        ("b", C.POINTER(C.c_float) * ORC_MAX_LAYERS),
        ("has_elu", C.c_int * ORC_MAX_LAYERS),
        ("elu_alpha", C.c_float * ORC_MAX_LAYERS),
        ("input_name", C.c_char * 64),
        ("output_name", C.c_char * 64),
        ("input_shape", C.c_int64 * 4),
        ("output_shape", C.c_int64 * 4),
        ("input_rank", C.c_int),
        ("output_rank", C.c_int),
    ]


class RawState(C.Structure):
    _fields_ = [
        ("quat", C.c_float * 4), ("gyro", C.c_float * 3), ("q", C.c_float * 12), ("dq", C.c_float * 12),
        ("axes", C.c_float * 4), ("foot_force", C.c_int16 * 4), ("joy_valid", C.c_int32), ("button0", C.c_int32),
    ]


class CtrlState(C.Structure):
    _fields_ = [
        ("H", C.c_int), ("vel_cmd", C.c_float * 3),
        ("g_hist", C.c_float * (3 * ORC_MAX_HIST)), ("w_hist", C.c_float * (3 * ORC_MAX_HIST)),
        ("cmd_hist", C.c_float * (3 * ORC_MAX_HIST)), ("q_hist", C.c_float * (12 * ORC_MAX_HIST)),
        ("dq_hist", C.c_float * (12 * ORC_MAX_HIST)), ("a_hist", C.c_float * (12 * ORC_MAX_HIST)),
        ("c_hist", C.c_uint16 * (4 * ORC_MAX_HIST)), ("action", C.c_float * 12),
        ("kp", C.c_float), ("kd", C.c_float),
    ]


class StepOut(C.Structure):
    _fields_ = [
        ("obs", C.c_float * (ORC_FRAME * ORC_MAX_HIST)), ("action_raw", C.c_float * 12),
        ("action", C.c_float * 12), ("q_des", C.c_double * 12), ("kp", C.c_double), ("kd", C.c_double),
    ]


def build(force: bool = False) -> None:
    """Compile the C restatement and, where the reference tree is present, oracle/_ref."""
    # make is incremental: always ask it, so an edited source never meets a stale library
    subprocess.run(["make", "-C", HERE, LIB_PATH] + (["-B"] if force else []), check=True, capture_output=True)
    subprocess.run(["make", "-C", HERE, "ref"], check=False, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(OrcModel)), C.c_char_p, C.c_int]
        L.orc_load.restype = C.c_int
        L.orc_free.argtypes = [C.POINTER(OrcModel)]
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_forward_f32.argtypes = [C.POINTER(OrcModel), fp, fp]
        L.orc_forward_f64.argtypes = [C.POINTER(OrcModel), fp, dp]
        L.orc_forward_rows_f32.argtypes = [C.POINTER(OrcModel), fp, fp, C.c_int64, C.c_int]
        L.orc_forward_blocked_f32.argtypes = [C.POINTER(OrcModel), fp, fp, C.c_int64, C.c_int]
        L.orc_forward_rows_f64.argtypes = [C.POINTER(OrcModel), fp, dp, C.c_int64, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_ctrl_reset.argtypes = [C.POINTER(CtrlState), C.c_int]
        L.orc_ctrl_assemble.argtypes = [C.POINTER(CtrlState), C.POINTER(RawState), fp]
        L.orc_ctrl_post.argtypes = [C.POINTER(CtrlState), C.POINTER(RawState), fp, C.POINTER(StepOut)]
        L.orc_ctrl_step.argtypes = [C.POINTER(CtrlState), C.POINTER(OrcModel), C.POINTER(RawState), C.c_int, C.POINTER(StepOut)]
        L.orc_ctrl_closed_loop_ns.argtypes = [C.POINTER(OrcModel), C.c_int, C.POINTER(RawState), C.c_int, C.c_int64,
                                              C.POINTER(C.c_uint64)]
        L.orc_ctrl_closed_loop_ns.restype = C.c_float
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CModel:
    def __init__(self, path: str):
        self._p = C.POINTER(OrcModel)()
        err = C.create_string_buffer(256)
        rc = lib().orc_load(path.encode(), C.byref(self._p), err, 256)
        if rc != 0:
            raise ValueError(err.value.decode())
        m = self._p.contents
        self.dims = [m.dims[i] for i in range(m.n_layers + 1)]
        self.in_dim, self.out_dim = self.dims[0], self.dims[-1]
        self.input_name = m.input_name.decode()
        self.output_name = m.output_name.decode()

    def __del__(self):
        try:
            if self._p:
                lib().orc_free(self._p)
        except Exception:
            pass

    def forward_f32(self, X: np.ndarray, threads: int = 1, blocked: bool = False) -> np.ndarray:
        X = np.ascontiguousarray(X, np.float32).reshape(-1, self.in_dim)
        Y = np.empty((X.shape[0], self.out_dim), np.float32)
        fn = lib().orc_forward_blocked_f32 if blocked else lib().orc_forward_rows_f32
        fn(self._p, _fp(X), _fp(Y), X.shape[0], threads)
        return Y

    def forward_f64(self, X: np.ndarray, threads: int = 1) -> np.ndarray:
        X = np.ascontiguousarray(X, np.float32).reshape(-1, self.in_dim)
        Y = np.empty((X.shape[0], self.out_dim), np.float64)
        lib().orc_forward_rows_f64(self._p, _fp(X), _dp(Y), X.shape[0], threads)
        return Y


def closed_loop_latency_ns(model: "CModel", raws, steps: int, H: int = 2) -> np.ndarray:
    """Per-step wall time (ns) of the restated publish() on ONE host thread, fp32 forward."""
    arr = (RawState * len(raws))(*raws)
    out = np.zeros(steps, np.uint64)
    lib().orc_ctrl_closed_loop_ns(model._p, H, arr, len(raws), steps, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def raw_from_py(r) -> RawState:
    """oracle.oracle.RawState -> C struct."""
    c = RawState()
    c.quat[:] = [float(v) for v in r.quat]
    c.gyro[:] = [float(v) for v in r.gyro]
    c.q[:] = [float(v) for v in r.q]
    c.dq[:] = [float(v) for v in r.dq]
    c.axes[:] = [float(v) for v in r.axes]
    c.foot_force[:] = [int(v) for v in r.foot_force]
    c.joy_valid = int(r.joy_valid)
    c.button0 = int(r.button0)
    return c


class CController:
    def __init__(self, model: CModel, H: int = 2):
        self.m = model
        self.s = CtrlState()
        self.H = H
        lib().orc_ctrl_reset(C.byref(self.s), H)

    def reset(self):
        lib().orc_ctrl_reset(C.byref(self.s), self.H)

    def step(self, raw: RawState, use_f64: bool = True) -> StepOut:
        out = StepOut()
        lib().orc_ctrl_step(C.byref(self.s), self.m._p, C.byref(raw), int(use_f64), C.byref(out))
        return out

    def assemble(self, raw: RawState) -> np.ndarray:
        obs = np.zeros(ORC_FRAME * self.H, np.float32)
        lib().orc_ctrl_assemble(C.byref(self.s), C.byref(raw), _fp(obs))
        return obs

    def post(self, raw: RawState, action_raw: np.ndarray) -> StepOut:
        out = StepOut()
        a = np.ascontiguousarray(action_raw, np.float32)
        lib().orc_ctrl_post(C.byref(self.s), C.byref(raw), _fp(a), C.byref(out))
        return out


class RefController:
    """The reference's own ONNXController (compiled from /root/reference, oracle/_ref), driven through its node
    interface.  The model it loads is the reference's onnx_inference/data/model.onnx (GO2_REF_ROOT, default
    /root/reference; the GPU box has no reference tree, so this class is only usable in the build container)."""

    def __init__(self):
        if not os.path.exists(REF_LIB_PATH):
            raise FileNotFoundError(REF_LIB_PATH)
        L = C.CDLL(REF_LIB_PATH)
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.refc_create.restype = C.c_void_p
        L.refc_destroy.argtypes = [C.c_void_p]
        L.refc_lowstate.argtypes = [C.c_void_p, fp, fp, fp, C.POINTER(C.c_int16)]
        L.refc_joy.argtypes = [C.c_void_p, fp, C.c_int, C.POINTER(C.c_int32), C.c_int]
        L.refc_robot.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.refc_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.refc_set_param.restype = C.c_int
        L.refc_step.argtypes = [C.c_void_p, fp, fp, dp, dp, dp]
        L.refc_step.restype = C.c_int
        self.L = L
        self.h = L.refc_create()
        if not self.h:
            raise RuntimeError("reference ONNXController could not be constructed")

    def close(self):
        if self.h:
            self.L.refc_destroy(self.h)
            self.h = None

    def feed(self, quat, gyro, foot_force, q, dq, axes, buttons, ready=True, safe=True):
        """One /lowstate + /joy delivery and the robot's joint state (what publish() reads)."""
        qf = np.ascontiguousarray(quat, np.float32); gf = np.ascontiguousarray(gyro, np.float32)
        ff = np.ascontiguousarray(foot_force, np.int16); acc = np.zeros(3, np.float32)
        self.L.refc_lowstate(self.h, _fp(qf), _fp(gf), _fp(acc), ff.ctypes.data_as(C.POINTER(C.c_int16)))
        ax = np.ascontiguousarray(axes, np.float32); bt = np.ascontiguousarray(buttons, np.int32)
        self.L.refc_joy(self.h, _fp(ax), ax.size, bt.ctypes.data_as(C.POINTER(C.c_int32)), bt.size)
        qd = np.ascontiguousarray(q, np.float64); dqd = np.ascontiguousarray(dq, np.float64)
        self.L.refc_robot(self.h, _dp(qd), _dp(dqd), int(ready), int(safe))

    def step(self):
        """Fires the 50 Hz timer.  Returns None if publish() was gated off, else (obs, action, q_des, kp, kd)."""
        obs = np.zeros(98, np.float32); act = np.zeros(12, np.float32)
        qd = np.zeros(12, np.float64); kp = np.zeros(12, np.float64); kd = np.zeros(12, np.float64)
        if not self.L.refc_step(self.h, _fp(obs), _fp(act), _dp(qd), _dp(kp), _dp(kd)):
            return None
        return obs, act, qd, kp, kd

    def set_param(self, name: str, value: float) -> bool:
        return bool(self.L.refc_set_param(self.h, name.encode(), float(value)))
