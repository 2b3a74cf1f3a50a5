"""Host-side mirror of the reference's operator interface for the hot path.

  ONNXActor      same surface as the reference class (reference: onnx_inference/include/onnx_actor.hpp:14-53,
                 usage as in onnx_inference/src/python/main.py:5-28): bind an observation and an action buffer,
                 act() fills the action in place.
  Go2Controller  the arithmetic of ONNXController::publish() (reference: onnx_controller/src/controller.cpp:173-251)
                 as one fused device step.
  PolicyBatch    many robots / rollouts through the same policy (device or host buffers).

Everything computes on the GPU through include/go2policy.h; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

DEFAULT_MODEL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "model.onnx")


def default_config(**overrides) -> capi.Config:
    cfg = capi.Config()
    capi.load().go2p_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if k == "q0":
            cfg.q0[:] = [float(x) for x in v]
        else:
            setattr(cfg, k, v)
    return cfg


class _Handle:
    def __init__(self, model_path: str, cfg: capi.Config | None = None):
        self.lib = capi.load()
        self.h = C.c_void_p()
        cfg = cfg if cfg is not None else default_config()
        capi.check(self.lib.go2p_create(os.fspath(model_path).encode(), C.byref(cfg), C.byref(self.h)))
        self.cfg = cfg
        info = capi.ModelInfo()
        capi.check(self.lib.go2p_model_info(self.h, C.byref(info)))
        self.info = info
        self.in_dim, self.out_dim = info.in_dim, info.out_dim
        self.input_name = info.input_name.decode()
        self.output_name = info.output_name.decode()
        self.dims = [info.dims[i] for i in range(info.n_layers + 1)]

    def close(self):
        if self.h:
            self.lib.go2p_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ONNXActor:
    """Python face of the reference's ONNXActor: observation/action numpy buffers are bound at construction
    (they must be C-contiguous float32 and outlive the actor) and act() overwrites ``action`` in place."""

    def __init__(self, model_path: str, observation: np.ndarray, action: np.ndarray, log_level: int = 2,
                 b1_mode: int | None = None, device: int = 0):
        for name, a in (("observation", observation), ("action", action)):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable):
                raise TypeError(f"{name} must be a writable C-contiguous float32 numpy array")
        kw = {"log_level": log_level, "device": device}
        if b1_mode is not None:
            kw["b1_mode"] = b1_mode
        self._hd = _Handle(model_path, default_config(**kw))
        self.observation, self.action = observation, action
        fp = C.POINTER(C.c_float)
        capi.check(self._hd.lib.go2p_bind(self._hd.h, observation.ctypes.data_as(fp), observation.size,
                                          action.ctypes.data_as(fp), action.size))

    def act(self) -> None:
        capi.check(self._hd.lib.go2p_act(self._hd.h))

    def check_dims(self) -> bool:
        return self.observation.size == self._hd.in_dim and self.action.size == self._hd.out_dim

    def print_model_info(self) -> None:
        print(f"Input dimension: {self._hd.in_dim}")
        print(f"Output dimension: {self._hd.out_dim}")
        print(f"Input name: {self._hd.input_name}")
        print(f"Output name: {self._hd.output_name}")

    def stats(self, reset: bool = False) -> capi.B1Stats:
        s = capi.B1Stats()
        capi.check(self._hd.lib.go2p_b1_stats_get(self._hd.h, C.byref(s), int(reset)))
        return s

    def close(self) -> None:
        self._hd.close()


class Go2Controller:
    """publish()'s arithmetic as one device step: raw robot state in, ObservationAction + joint targets out."""

    def __init__(self, model_path: str = DEFAULT_MODEL, **cfg):
        self._hd = _Handle(model_path, default_config(**cfg))

    def step(self, raw: capi.RawState) -> capi.StepOut:
        out = capi.StepOut()
        capi.check(self._hd.lib.go2p_step_fused(self._hd.h, C.byref(raw), C.byref(out)))
        return out

    def step_cmd(self, raw: capi.RawState):
        """step() plus the same step's send_command arguments in Unitree motor order (controller.cpp:235-251)."""
        out, cmd = capi.StepOut(), capi.MotorCmd()
        capi.check(self._hd.lib.go2p_step_fused_cmd(self._hd.h, C.byref(raw), C.byref(out), C.byref(cmd)))
        return out, cmd

    def log_enable(self, capacity: int) -> None:
        """ObservationAction ring on the device (ObservationAction.msg:1-2); 0 turns it off."""
        capi.check(self._hd.lib.go2p_log_enable(self._hd.h, int(capacity)))

    def log_drain(self, max_records: int = 4096):
        """Records written since the last drain, oldest first: ([n, in_dim] observations, [n, 12] actions, dropped)."""
        rec = self._hd.in_dim + 12
        buf = np.zeros((max_records, rec), np.float32)
        n, dropped = C.c_int(), C.c_uint64()
        capi.check(self._hd.lib.go2p_log_drain(self._hd.h, buf.ctypes.data, max_records, C.byref(n), C.byref(dropped)))
        buf = buf[: n.value]
        return buf[:, : self._hd.in_dim].copy(), buf[:, self._hd.in_dim:].copy(), int(dropped.value)

    def closed_loop(self, raws, steps: int):
        """`steps` fused steps over the raw states (cycled), timed in native code.
        Returns (host_ns[steps], device_ns[steps], last StepOut)."""
        arr = (capi.RawState * len(raws))(*raws)
        host = np.zeros(steps, np.uint64)
        dev = np.zeros(steps, np.uint64)
        last = capi.StepOut()
        u64 = C.POINTER(C.c_uint64)
        capi.check(self._hd.lib.go2p_b1_closed_loop(self._hd.h, arr, len(raws), steps, host.ctypes.data_as(u64),
                                                    dev.ctypes.data_as(u64), C.byref(last)))
        return host, dev, last

    def selfdriven(self, raws, steps: int):
        """Profiling twin of the resident kernel: one bounded launch of `steps` closed-loop steps.
        Returns (last published action[12], elapsed ms)."""
        arr = (capi.RawState * len(raws))(*raws)
        act = np.zeros(12, np.float32)
        ms = C.c_float()
        capi.check(self._hd.lib.go2p_b1_selfdriven(self._hd.h, arr, len(raws), steps,
                                                   act.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ms)))
        return act, float(ms.value)

    def reset(self) -> None:
        capi.check(self._hd.lib.go2p_reset_history(self._hd.h))

    def set_gains(self, kp: float, kd: float) -> None:
        capi.check(self._hd.lib.go2p_set_gains(self._hd.h, kp, kd))

    def stop(self) -> None:
        capi.check(self._hd.lib.go2p_persistent_stop(self._hd.h))

    def stats(self, reset: bool = False) -> capi.B1Stats:
        s = capi.B1Stats()
        capi.check(self._hd.lib.go2p_b1_stats_get(self._hd.h, C.byref(s), int(reset)))
        return s

    def close(self) -> None:
        self._hd.close()


class PolicyBatch:
    """Batched inference.  Device tensors are passed as raw pointers (torch ``.data_ptr()`` or go2p_dev_alloc)."""

    def __init__(self, model_path: str = DEFAULT_MODEL, device: int = 0, **cfg):
        self._hd = _Handle(model_path, default_config(device=device, **cfg))
        self.in_dim, self.out_dim = self._hd.in_dim, self._hd.out_dim

    @property
    def info(self) -> capi.ModelInfo:
        return self._hd.info

    def infer_device(self, d_obs: int, d_act: int, B: int, precision: int = capi.PREC_FP32, stream: int = 0,
                     d_button0: int | None = None, d_qdes: int | None = None, flags: int = 0, d_cmd: int | None = None) -> None:
        """precision defaults to the reference's (fp32, 1e-5 contract); the tensor-core precisions are opt-in."""
        capi.check(self._hd.lib.go2p_infer_batch_cmd(self._hd.h, d_obs, d_button0, d_act, d_qdes, d_cmd, B, precision, flags,
                                                     stream or None))

    def time_device(self, d_obs: int, d_act: int, B: int, precision: int, iters: int, stream: int = 0,
                    d_button0: int | None = None, d_qdes: int | None = None, flags: int = 0) -> float:
        ms = C.c_float()
        capi.check(self._hd.lib.go2p_time_batch(self._hd.h, d_obs, d_button0, d_act, d_qdes, B, precision, flags,
                                                stream or None, iters, C.byref(ms)))
        return float(ms.value)

    def infer_host(self, obs: np.ndarray, act: np.ndarray | None = None, precision: int = capi.PREC_FP32) -> np.ndarray:
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, self.in_dim)
        if act is None:
            act = np.empty((obs.shape[0], self.out_dim), np.float32)
        capi.check(self._hd.lib.go2p_infer_batch_host(self._hd.h, obs.ctypes.data, act.ctypes.data, obs.shape[0], precision))
        return act

    def saturation_count(self, reset: bool = True, stream: int = 0) -> int:
        """(row, operand block) pairs clipped at +-65504 by fp16 launches made with F_SAT_COUNT since the last reset."""
        n = C.c_uint64()
        capi.check(self._hd.lib.go2p_saturation_count(self._hd.h, C.byref(n), int(reset), stream or None))
        return int(n.value)

    def last_launches(self) -> int:
        return int(self._hd.lib.go2p_last_launch_count(self._hd.h))

    def pinned(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy view over cudaHostAlloc'ed memory (freed with the handle's process)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        capi.check(self._hd.lib.go2p_host_alloc(self._hd.h, n, C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def assemble_device(self, d_raw: int, d_prev_action: int | None, d_vel_cmd: int, d_obs: int, B: int, stream: int = 0):
        capi.check(self._hd.lib.go2p_assemble_batch(self._hd.h, d_raw, d_prev_action, d_vel_cmd, d_obs, B, stream or None))

    def step_device(self, d_raw: int, d_vel_cmd: int, d_obs: int, d_action: int, d_qdes: int, B: int,
                    precision: int = capi.PREC_FP32, stream: int = 0) -> None:
        """Batched publish() (reference: controller.cpp:173-251 per robot): raw states -> observation history ->
        policy -> clamp/mask -> q_des, all on device buffers; d_action is the previous published action on entry."""
        capi.check(self._hd.lib.go2p_step_batch(self._hd.h, d_raw, d_vel_cmd, d_obs, d_action, d_qdes, B, precision,
                                                stream or None))

    def step_device_cmd(self, d_raw: int, d_vel_cmd: int, d_obs: int, d_action: int, d_qdes: int | None, d_cmd: int | None,
                        B: int, precision: int = capi.PREC_FP32, stream: int = 0) -> None:
        """step_device ending in motor commands: d_cmd [B] go2p_motor_cmd (Unitree motor order), d_qdes may be None."""
        capi.check(self._hd.lib.go2p_step_batch_cmd(self._hd.h, d_raw, d_vel_cmd, d_obs, d_action, d_qdes, d_cmd, B, precision,
                                                    stream or None))

    def step_host(self, raw: np.ndarray, action: np.ndarray, cmd: np.ndarray | None = None,
                  precision: int = capi.PREC_FP32) -> None:
        """Closed-loop step for len(raw) robots from host buffers; per-robot history stays on the device.
        raw: uint8 [B, 156] (go2p_raw_state records), action: float32 [B, 12] out, cmd: uint8 [B, 112] out or None."""
        B = raw.shape[0]
        capi.check(self._hd.lib.go2p_step_batch_host(self._hd.h, raw.ctypes.data, action.ctypes.data,
                                                     cmd.ctypes.data if cmd is not None else None, B, precision))

    def step_host_reset(self) -> None:
        capi.check(self._hd.lib.go2p_step_batch_host_reset(self._hd.h))

    def close(self) -> None:
        self._hd.close()


class Fleet:
    """One process, several GPUs: rows are split into contiguous shards, one handle + host thread per device
    (include/go2policy.h: go2p_fleet_*).  No data crosses between devices."""

    def __init__(self, model_path: str = DEFAULT_MODEL, devices=(0,), **cfg):
        self.lib = capi.load()
        self.f = C.c_void_p()
        c = default_config(**cfg)
        dev = (C.c_int32 * len(devices))(*devices)
        capi.check(self.lib.go2p_fleet_create(os.fspath(model_path).encode(), C.byref(c), dev, len(devices), C.byref(self.f)))
        self.n_devices = len(devices)

    def infer_host(self, obs: np.ndarray, act: np.ndarray, precision: int = capi.PREC_FP32) -> None:
        capi.check(self.lib.go2p_fleet_infer_host(self.f, obs.ctypes.data, act.ctypes.data, obs.shape[0], precision))

    def step_host(self, raw: np.ndarray, action: np.ndarray, cmd: np.ndarray | None = None, precision: int = capi.PREC_FP32) -> None:
        capi.check(self.lib.go2p_fleet_step_host(self.f, raw.ctypes.data, action.ctypes.data,
                                                 cmd.ctypes.data if cmd is not None else None, raw.shape[0], precision))

    def close(self) -> None:
        if self.f:
            self.lib.go2p_fleet_destroy(self.f)
            self.f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_rows(total_rows: int, n: int, i: int):
    """(begin, end) of shard i of n as the C ABI splits them (go2p_shard_rows)."""
    b, e = C.c_int64(), C.c_int64()
    capi.check(capi.load().go2p_shard_rows(total_rows, n, i, C.byref(b), C.byref(e)))
    return b.value, e.value
