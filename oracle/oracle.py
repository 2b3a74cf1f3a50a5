"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

numpy restatement of the reference hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
leg may import this module.

PARITY STATUS
  * A7 (the policy forward pass) -- "parity unpinned": the arithmetic of the
    reference lives in the un-vendored onnxruntime 1.20.1 prebuilt binary
    (reference: onnx_inference/cmake/dependencies.cmake:17) which is not
    available in this image, and the reference ships no golden vector.  This
    restatement follows the ONNX opset-17 operator definitions of the graph
    stored in onnx_inference/data/model.onnx and is cross-checked against an
    independent torch-CPU fp64 evaluation (tests/golden/make_golden.py).
  * A1-A6, A9-A11 (observation assembly / action post-processing) -- pinned against the
    reference's own controller.cpp compiled from where it lies (oracle/Makefile ->
    oracle/_ref/libref_controller.so; external ROS / Eigen / ORT headers replaced by the
    stubs under oracle/ref_stubs/): bit-identical over the closed-loop fixture
    (tests/test_ref_controller.py).  The stub's quaternion arithmetic restates Eigen 3.4.

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .onnx_mini import Policy, load_policy  # noqa: F401  (re-export)

F32 = np.float32
F64 = np.float64

# reference: onnx_controller/include/onnx_controller/controller.hpp:13-16
K_DIM_DOF = 12
K_DIM_OBS = 49
K_HISTORY = 2
K_ACTION_LIMIT = F32(1000.0)
# reference: controller.hpp:165 (Isaac joint order, double)
Q0 = np.array([0.1, -0.1, 0.1, -0.1, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5], dtype=F64)
# reference: controller.hpp:100-103 (unitree <-> isaac foot order, threshold 22)
FOOT_PERM = (1, 0, 3, 2)
FOOT_THRESHOLD = 22
# reference: controller.cpp:244 / :246 / controller.hpp:119-120
ACTION_SCALE = 0.25
KP_DEFAULT = F32(28.0)
KD_DEFAULT = F32(0.5)
KP_DEADMAN = 5
# term widths of one observation frame, in concat order (controller.cpp:210-212)
TERM_WIDTHS = (3, 3, 3, 12, 12, 12, 4)


# ----------------------------------------------------------------------------
# A7: policy forward  (reference: onnx_actor.cpp:38-48 -> Ort::Session::Run)
# ----------------------------------------------------------------------------

def elu(x: np.ndarray, alpha: float) -> np.ndarray:
    """ONNX Elu-6: f(x) = alpha*(exp(x)-1) for x < 0, x otherwise.
    NaN and -0.0 take the 'otherwise' branch and pass through unchanged."""
    with np.errstate(over="ignore", invalid="ignore"):
        neg = x < 0
        e = np.exp(np.where(neg, x, 0)) - 1
        return np.where(neg, x.dtype.type(alpha) * e, x)


def forward(policy: Policy, obs: np.ndarray, dtype=F64) -> np.ndarray:
    """Gemm(alpha=1,beta=1,transB=1) / Elu chain of the parsed graph.
    obs is [B, in] or [in]; fp32 inputs/weights are widened to ``dtype``."""
    x = np.asarray(obs, dtype=F32).astype(dtype)
    squeeze = x.ndim == 1
    if squeeze:
        x = x[None, :]
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for op, c in getattr(policy, "pre_ops", ()):        # observation normaliser nodes, as written in the graph
            c = c.astype(dtype)
            x = x - c if op == "Sub" else x + c if op == "Add" else x * c if op == "Mul" else x / c
        for layer in policy.layers:
            x = x @ layer.weight.astype(dtype).T + layer.bias.astype(dtype)
            if layer.elu_alpha is not None:
                x = np.maximum(x, 0) if layer.elu_alpha == 0.0 else elu(x, layer.elu_alpha)
    return x[0] if squeeze else x


def forward_operand_rounded(policy: Policy, obs: np.ndarray, rounder) -> np.ndarray:
    """fp64 forward where every Gemm operand (activation and weight) is first
    rounded by ``rounder`` (bf16 / fp16 / tf32) -- models a tensor-core path
    with fp32 accumulation; used to derive the stated tolerances."""
    x = np.asarray(obs, dtype=F32)
    if x.ndim == 1:
        x = x[None, :]
    for layer in policy.layers:
        xr = rounder(x.astype(F32)).astype(F64)
        wr = rounder(layer.weight).astype(F64)
        y = xr @ wr.T + layer.bias.astype(F64)
        if layer.elu_alpha is not None:
            y = elu(y, layer.elu_alpha)
        x = y.astype(F32)
    return x.astype(F64)


def round_bf16(a: np.ndarray) -> np.ndarray:
    """round-to-nearest-even fp32 -> bf16 -> fp32 (finite inputs)."""
    u = np.ascontiguousarray(a, dtype=F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(F32)


def round_fp16(a: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return np.asarray(a, dtype=F32).astype(np.float16).astype(F32)


def round_tf32(a: np.ndarray) -> np.ndarray:
    """round-to-nearest-even to 10 explicit mantissa bits."""
    u = np.ascontiguousarray(a, dtype=F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000).astype(np.uint32)
    return r.view(F32)


# ----------------------------------------------------------------------------
# A9 / A11: action post-processing  (reference: controller.cpp:217-223,235-248)
# ----------------------------------------------------------------------------

def clamp_mask(action: np.ndarray, button0) -> np.ndarray:
    """std::clamp(a,-1000,1000) == (a<lo)?lo:(hi<a)?hi:a  -- NaN passes through;
    then a *= (buttons[0]==0) as a float multiply (keeps sign of zero, NaN)."""
    a = np.asarray(action, dtype=F32)
    lo, hi = -K_ACTION_LIMIT, K_ACTION_LIMIT
    with np.errstate(invalid="ignore"):
        c = np.where(a < lo, lo, np.where(hi < a, hi, a)).astype(F32)
        m = (np.asarray(button0) == 0).astype(F32)
        if m.ndim == 1 and c.ndim == 2:
            m = m[:, None]
        return (c * m).astype(F32)


def joint_targets(action: np.ndarray, button0, kp=KP_DEFAULT, kd=KD_DEFAULT):
    """q_des = q0 + (double)a * 0.25; kp = button0==0 ? kp_ : 5; kd = kd_
    (reference: controller.cpp:244-247)."""
    a = np.asarray(action, dtype=F32).astype(F64)
    q_des = Q0 + a * ACTION_SCALE
    kp_eff = np.where(np.asarray(button0) == 0, F64(F32(kp)), F64(KP_DEADMAN))
    return q_des, kp_eff, F64(F32(kd))


# Isaac joint order (reference: controller.hpp:168-170, breadth-first: joint*4 + leg, legs FL FR RL RR) -> Unitree motor
# order (SDK: leg*3 + joint, legs FR FL RR RL).  The permutation itself lives in the out-of-tree Go2RobotInterface
# (README.md:5); it is the same left/right swap as the foot contacts, controller.hpp:100-103.
ISAAC_JOINTS = [f"{leg}_{j}" for j in ("hip", "thigh", "calf") for leg in ("FL", "FR", "RL", "RR")]
UNITREE_MOTORS = [f"{leg}_{j}" for leg in ("FR", "FL", "RR", "RL") for j in ("hip", "thigh", "calf")]
ISAAC_OF_MOTOR = np.array([ISAAC_JOINTS.index(n) for n in UNITREE_MOTORS])


def motor_command(q_des_isaac: np.ndarray, kp, kd):
    """send_command arguments (controller.cpp:235-251) re-ordered for the motors: (q_des[..., 12] motor order, kp, kd)."""
    return np.asarray(q_des_isaac)[..., ISAAC_OF_MOTOR], kp, kd


# ----------------------------------------------------------------------------
# A1-A6: observation assembly  (reference: controller.cpp:173-212)
# ----------------------------------------------------------------------------

def vel_cmd_from_axes(axes) -> np.ndarray:
    """reference: controller.cpp:176-178.  axes[0]**2 goes through double pow,
    sign branch is (>0 ? 1 : -1) so axes[0]==0 yields -0.0f."""
    ax = np.asarray(axes, dtype=F32)
    v0 = ax[1]
    sgn = 1 if ax[0] > 0 else -1
    v1 = F32((F64(ax[0]) * F64(ax[0])) * sgn * 0.8)
    v2 = F32(ax[3] * ax[1])
    return np.array([v0, v1, v2], dtype=F32)


def gravity_body(quat_wxyz) -> np.ndarray:
    """reference: controller.cpp:182-184 -- Eigen: quaternion_.inverse() * (0,0,-1).
    inverse() = conj/squaredNorm (zero quaternion if squaredNorm == 0);
    q*v = v + w*(2 u x v) + u x (2 u x v); all fp32, no FMA contraction.
    squaredNorm is summed as ((x^2+y^2)+z^2)+w^2 (Eigen's SIMD reduction order
    is unspecified -> tolerance of a few ulp vs. a real Eigen build)."""
    w, x, y, z = (F32(v) for v in quat_wxyz)
    n2 = F32(F32(F32(x * x) + F32(y * y)) + F32(z * z)) + F32(w * w)
    n2 = F32(n2)
    if n2 > 0:
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            cw, cx, cy, cz = F32(w / n2), F32(-x / n2), F32(-y / n2), F32(-z / n2)
    else:
        cw = cx = cy = cz = F32(0)
    v = (F32(0), F32(0), F32(-1))
    u = (cx, cy, cz)

    def cross(a, b):
        return (F32(F32(a[1] * b[2]) - F32(a[2] * b[1])),
                F32(F32(a[2] * b[0]) - F32(a[0] * b[2])),
                F32(F32(a[0] * b[1]) - F32(a[1] * b[0])))

    with np.errstate(over="ignore", invalid="ignore"):
        uv = cross(u, v)
        uv = tuple(F32(c + c) for c in uv)
        uuv = cross(u, uv)
        out = [F32(F32(v[i] + F32(cw * uv[i])) + uuv[i]) for i in range(3)]
    return np.array(out, dtype=F32)


def contacts_from_foot_force(foot_force) -> np.ndarray:
    """reference: controller.hpp:99-103."""
    ff = np.asarray(foot_force, dtype=np.int16)
    return np.array([1 if ff[p] >= FOOT_THRESHOLD else 0 for p in FOOT_PERM], dtype=np.uint16)


@dataclass
class RawState:
    """Everything publish() reads from the outside world in one control step."""
    quat: np.ndarray            # float32[4]  (w,x,y,z)  controller.hpp:95-97
    gyro: np.ndarray            # float32[3]  controller.hpp:109
    q: np.ndarray               # float32[12] (float)get_q()[i]  controller.cpp:189
    dq: np.ndarray              # float32[12] controller.cpp:190
    foot_force: np.ndarray      # int16[4]    controller.hpp:100-103 (unitree order)
    axes: np.ndarray            # float32[4]  joy axes (0,1,3 are used)
    joy_valid: int = 1          # joy_ && !axes.empty()  controller.cpp:173
    button0: int = 0            # joy_->buttons[0]  controller.cpp:221,246


@dataclass
class ControllerState:
    """Member state of ONNXController that persists across steps
    (reference: controller.hpp:132-162); H generalises kHistory."""
    H: int = K_HISTORY
    vel_cmd: np.ndarray = field(default=None)
    hist: list = field(default=None)       # 7 term histories, widths TERM_WIDTHS*H
    action: np.ndarray = field(default=None)

    def __post_init__(self):
        self.reset()

    def reset(self):
        self.vel_cmd = np.zeros(3, F32)
        self.hist = [np.zeros(w * self.H, F32) for w in TERM_WIDTHS]
        self.action = np.zeros(K_DIM_DOF, F32)


def assemble_observation(state: ControllerState, raw: RawState) -> np.ndarray:
    """reference: controller.cpp:173-212 (+ populate_buffer controller.hpp:45-68).
    Mutates ``state`` exactly as publish() mutates the node's members and returns
    the 49*H observation fed to the policy."""
    if raw.joy_valid:
        state.vel_cmd = vel_cmd_from_axes(raw.axes)
    g = gravity_body(raw.quat)
    qf = np.asarray(raw.q, dtype=F32)
    q_rel = (qf.astype(F64) - Q0).astype(F32)          # controller.cpp:194-197 (double subtract)
    dq = np.asarray(raw.dq, dtype=F32)
    contacts = contacts_from_foot_force(raw.foot_force).astype(F32)
    cur = [g, np.asarray(raw.gyro, F32), state.vel_cmd, q_rel, dq, state.action, contacts]
    for h, c in zip(state.hist, cur):                  # shift-left by n, append
        n = c.size
        h[:-n] = h[n:]
        h[-n:] = c
    return np.concatenate(state.hist).astype(F32)


@dataclass
class StepOut:
    obs: np.ndarray
    action_raw: np.ndarray   # policy output before clamp/mask
    action: np.ndarray       # published action (post clamp+mask)
    q_des: np.ndarray        # double[12]
    kp: float
    kd: float


def controller_step(policy: Policy, state: ControllerState, raw: RawState,
                    dtype=F64, kp=KP_DEFAULT, kd=KD_DEFAULT, act_fn=None) -> StepOut:
    """One full publish() (reference: controller.cpp:173-251) minus ROS I/O.
    ``act_fn(obs)->action`` overrides the policy (used to feed a device action
    back for closed-loop comparisons)."""
    obs = assemble_observation(state, raw)
    a_raw = (act_fn(obs) if act_fn is not None else forward(policy, obs, dtype)).astype(F32)
    a = clamp_mask(a_raw, raw.button0)
    state.action = a.copy()
    q_des, kp_eff, kd_eff = joint_targets(a, raw.button0, kp, kd)
    return StepOut(obs, a_raw, a, q_des, float(kp_eff), float(kd_eff))


# ----------------------------------------------------------------------------
# synthetic input sets of SURVEY.md section 8(d)
# ----------------------------------------------------------------------------

def make_raw_states(n: int, seed: int = 2):
    """Config-2 closed-loop raw-state distribution."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis) + 1e-12
        ang = rng.normal(0.0, 0.2)
        quat = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis]).astype(F32)
        out.append(RawState(
            quat=quat,
            gyro=rng.normal(0, 0.5, 3).astype(F32),
            q=(Q0 + rng.normal(0, 0.3, 12)).astype(F32),
            dq=rng.normal(0, 2.0, 12).astype(F32),
            foot_force=rng.integers(0, 61, 4).astype(np.int16),
            axes=rng.uniform(-1, 1, 4).astype(F32),
            joy_valid=1,
            button0=int(rng.random() < 0.01),
        ))
    return out


def make_obs_d1(b: int, in_dim: int, seed: int = 0) -> np.ndarray:
    """D1: iid N(0,1) fp32 observations."""
    return np.random.default_rng(seed).standard_normal((b, in_dim)).astype(F32)


def make_obs_d2(policy: Policy, b: int, seed: int = 1, H: int = K_HISTORY) -> np.ndarray:
    """D2: realistic observations -- each row is the obs of a 3-step open-loop
    rollout from reset with the config-2 raw-state distribution."""
    raws = make_raw_states(3 * b, seed)
    rows = np.zeros((b, K_DIM_OBS * H), F32)
    for i in range(b):
        st = ControllerState(H=H)
        for s in range(3):
            r = raws[3 * i + s]
            obs = assemble_observation(st, r)
            st.action = np.random.default_rng(seed * 7919 + 3 * i + s).normal(0, 1.5, 12).astype(F32)
        rows[i] = obs
    return rows
