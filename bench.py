#!/usr/bin/env python
"""Headline benchmark of the Go2 policy hot path (BASELINE.json: policy inferences/s, batch-1 latency).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the batched hot path (A7 with the fused A9 clamp+mask epilogue) over one batch of synthetic
observations: BASELINE.json configs[3], 1,048,576 rollouts of the bundled Go2 policy PER GPU (weak scaling: rows are
independent, every rank owns its block and its own copy of the weights, no collective on the data path).
`value`   device-timed throughput, observations already resident in HBM (411 MB per step per GPU > 126 MB L2).
`e2e`     the same metric through go2p_infer_batch_host with pinned HOST buffers (H2D + kernel + D2H inside).
`roofline` algorithmic 440 B/inference over the kernel's measured launch duration vs measured HBM.
`cpu_baseline` the oracle's C port of the reference path on the box's host cores (ORT itself is not available).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_INF = 98 * 4 + 12 * 4          # obs in + action out, weights amortised
ALGO_FLOP_PER_INF = 93696
# dram__bytes_read.sum + dram__bytes_write.sum of one tc_mlp_kernel launch over 1,048,576 rows, from the ncu
# --set full capture summarised in profiles/r01_tc_mlp_kernel_final_details.txt (415.6 MB read + 44.1 MB
# written; part of the 50 MB of actions is still in L2 when the launch ends -- the capture before it showed 63 MB)
NCU_TRAFFIC_BYTES_PER_ROW = (415.561984e6 + 44.075008e6) / 1048576
METRIC = "policy_inferences_per_sec"
UNIT = "inferences/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_raw_states(capi, n, seed):
    """Closed-loop raw-state distribution of SURVEY.md 8(d) (small-angle attitude, joints around the default pose,
    uniform joystick, rare dead-man press) as an array of go2p_raw_state -- bench input only, generated here so the
    timed paths never touch oracle/."""
    import numpy as np
    rng = np.random.default_rng(seed)
    q0 = np.array([0.1, -0.1, 0.1, -0.1, 0.8, 0.8, 1.0, 1.0, -1.5, -1.5, -1.5, -1.5], np.float32)   # controller.hpp:119-120
    arr = (capi.RawState * n)()
    for i in range(n):
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis) + 1e-12
        ang = rng.normal(0.0, 0.2)
        r = arr[i]
        r.quat[:] = [float(np.cos(ang / 2))] + [float(v) for v in np.sin(ang / 2) * axis]
        r.gyro[:] = [float(v) for v in rng.normal(0, 0.5, 3)]
        r.q[:] = [float(v) for v in q0 + rng.normal(0, 0.3, 12)]
        r.dq[:] = [float(v) for v in rng.normal(0, 2.0, 12)]
        r.foot_force[:] = [int(v) for v in rng.integers(0, 61, 4)]
        r.axes[:] = [float(v) for v in rng.uniform(-1, 1, 4)]
        r.joy_valid = 1
        r.button0 = int(rng.random() < 0.01)
    return arr


def cpu_reference_rate(rows_per_step, steps, warmup, threads=None, blocked=False):
    """The reference path on host cores: B independent batch-1 forwards (the static-batch model's semantics,
    onnx_actor.cpp:38-48) + A9 clamp/mask, rows split over all cores -- C port in oracle/ (test infrastructure)."""
    import numpy as np
    from oracle import coracle, oracle
    from go2_onnx_controller_b200 import DEFAULT_MODEL
    coracle.build()
    cm = coracle.CModel(DEFAULT_MODEL)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)
    threads = threads or len(os.sched_getaffinity(0))
    X = oracle.make_obs_d1(rows_per_step, 98, seed=0)
    for _ in range(warmup):
        cm.forward_f32(X[: max(1, rows_per_step // 8)], threads, blocked)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = cm.forward_f32(X, threads, blocked)
        oracle.clamp_mask(y, 0)
    dt = time.perf_counter() - t0
    return rows_per_step * steps / dt, dt, threads


def run_reference(args, rank):
    if rank != 0:
        return
    rows = args.ref_rows
    rate, dt, threads = cpu_reference_rate(rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: bundled Go2 policy 98-128-128-128-12, batched rollouts; reference arm = CPU, "
                               f"{rows} rows per step (bounded sample of the 1,048,576-row step)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{rows} rows x {args.steps} steps, C restatement of Gemm/Elu chain + clamp/mask, "
                                   "batch-1 semantics per row, OpenMP over all host cores; ONNX Runtime itself is not "
                                   "installable here (no wheel, no network)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_048_576, help="rows per GPU per step")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--ref-rows", type=int, default=262_144)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--b1-steps", type=int, default=100_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-b1", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import go2_onnx_controller_b200 as pkg
    from go2_onnx_controller_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prec = capi.PREC_NAMES[args.precision]
    rows = args.rows
    pb = pkg.PolicyBatch(pkg.DEFAULT_MODEL, device=local_rank)
    g = torch.Generator(device="cuda").manual_seed(rank)
    d_obs = torch.randn((rows, 98), device="cuda", dtype=torch.float32, generator=g)     # D1, seed = rank
    d_act = torch.empty((rows, 12), device="cuda", dtype=torch.float32)
    d_b0 = torch.zeros((rows,), device="cuda", dtype=torch.int32)
    stream = torch.cuda.current_stream().cuda_stream
    flags = capi.F_CLAMP_MASK

    def step():
        pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), rows, prec, stream, d_b0.data_ptr(), None, flags)

    for _ in range(args.warmup):
        step()
    launches_per_step = pb.last_launches()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    barrier()
    ms = max_over_ranks(ms_local)
    # keep the device busy a little longer for short runs so that the 100 ms clock sampler sees it under load
    if rank == 0 and time.time() - t_wall0 < 0.6:
        tb = time.time()
        while time.time() - tb < 0.6:
            step()
        torch.cuda.synchronize()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * rows * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps

    # parity spot check of what was just timed (oracle = checker only)
    if rank == 0:
        from oracle import oracle
        pol = oracle.load_policy(pkg.DEFAULT_MODEL)
        idx = torch.randint(0, rows, (256,), device="cuda")
        ref = oracle.clamp_mask(oracle.forward(pol, d_obs[idx].cpu().numpy()).astype(np.float32), 0)
        got = d_act[idx].cpu().numpy()
        parity_err = float(np.abs(got - ref).max())
    # ---- BASELINE.json configs[2]: batch 4096 on one GPU -- single-launch latency and pipelined throughput
    small = None
    if rank == 0:
        sb = 4096
        lat = []
        for _ in range(200):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pb.infer_device(d_obs.data_ptr(), d_act.data_ptr(), sb, prec, stream, d_b0.data_ptr(), None, flags)
            a1.record()
            a1.synchronize()
            lat.append(a0.elapsed_time(a1) * 1e3)
        ms_pipe = pb.time_device(d_obs.data_ptr(), d_act.data_ptr(), sb, prec, 200, stream, d_b0.data_ptr(), None, flags)
        small = {"batch": sb, "single_launch_us_p50": float(np.percentile(lat, 50)), "single_launch_us_p99": float(np.percentile(lat, 99)),
                 "pipelined_inferences_per_sec": sb * 200 / (ms_pipe * 1e-3), "note": "L2-resident (1.8 MB), launch-latency bound"}
    # ---- e2e: HOST buffers through the C ABI, copies inside the timed region
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    hx = pb.pinned((rows, 98))
    hy = pb.pinned((rows, 12))
    hx[:] = d_obs.cpu().numpy()
    for _ in range(2):
        pb.infer_host(hx, hy, prec)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pb.infer_host(hx, hy, prec)          # synchronous: returns when the actions are in host memory
    dt_local = time.perf_counter() - t0
    e2e_launches = pb.last_launches()
    barrier()
    dt = max_over_ranks(dt_local)
    e2e_value = world * rows * e2e_steps / dt

    hbm_peak, tf_peak, peak_src = measured_peaks()
    achieved_gbs = ALGO_BYTES_PER_INF * rows / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"configs[3]: bundled Go2 policy 98-128-128-128-12, {rows} rollouts per GPU per step "
                               f"(N(0,1) observations, seed=rank), A7 + fused A9 clamp/mask, {args.precision} operands / fp32 accumulate",
                   "rows_per_gpu": rows, "l2": "inputs larger than L2 (411 MB/step vs 126 MB), no flush needed",
                   "sharding": f"dp{world}, contiguous row blocks, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rows * 98 * 4, "d2h_bytes_per_step": rows * 12 * 4,
                "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3, "api": "go2p_infer_batch_host (pinned host buffers)"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                     "traffic": (NCU_TRAFFIC_BYTES_PER_ROW * rows if args.precision != "fp32" else None), "peak_source": peak_src, "kernel": "tc_mlp_kernel" if args.precision != "fp32" else "sgemm_bias_act_kernel",
                     "algorithmic_bytes_per_inference": ALGO_BYTES_PER_INF,
                     "tensor_frac": ALGO_FLOP_PER_INF * rows / (kernel_ms * 1e-3) / 1e12 / tf_peak},
        "clocks": clocks,
    }
    if rank == 0:
        line["parity_max_abs_err_vs_oracle"] = parity_err
        line["batch4096"] = small
    if rank == 0 and world == 1:
        # the whole publish() for `rows` robots (raw state -> history -> policy -> clamp/mask -> q_des),
        # two launches per step, per-robot state resident in HBM (go2p_step_batch)
        import ctypes as C
        arr = synthetic_raw_states(capi, 4096, seed=3)
        raw_np = np.frombuffer(bytes(arr), np.uint8).reshape(4096, C.sizeof(capi.RawState))
        d_raw = torch.from_numpy(np.tile(raw_np, (rows // 4096 + 1, 1))[:rows].copy()).to("cuda")
        s_obs = torch.zeros((rows, 98), device="cuda"); s_vel = torch.zeros((rows, 3), device="cuda")
        s_act = torch.zeros((rows, 12), device="cuda"); s_q = torch.zeros((rows, 12), device="cuda", dtype=torch.float64)
        def ctl_step():
            pb.step_device(d_raw.data_ptr(), s_vel.data_ptr(), s_obs.data_ptr(), s_act.data_ptr(), s_q.data_ptr(), rows, prec, stream)
        for _ in range(3):
            ctl_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_ctl = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n_ctl):
            ctl_step()
        e1.record(); torch.cuda.synchronize()
        ctl_ms = e0.elapsed_time(e1) / n_ctl
        line["batched_step"] = {"robots": rows, "ms_per_step": ctl_ms, "robot_steps_per_sec": rows / (ctl_ms * 1e-3),
                                "launches_per_step": pb.last_launches(),
                                "note": "go2p_step_batch: A1-A6 assembly kernel + fused A7/A9/A11 kernel, state in HBM"}
        del d_raw, s_obs, s_vel, s_act, s_q
    if rank == 0 and world == 1 and not args.no_b1:
        # BASELINE.json configs[1]: batch-1 closed loop, fused pre/post, resident kernel
        raws = list(synthetic_raw_states(capi, 512, seed=2))
        ctl = pkg.Go2Controller(pkg.DEFAULT_MODEL, device=local_rank)
        ctl.closed_loop(raws, 10_000)
        host_ns, dev_ns, _ = ctl.closed_loop(raws, args.b1_steps)
        ctl.close()
        line["b1_latency_us"] = {"steps": int(args.b1_steps), "p50": float(np.percentile(host_ns, 50)) / 1e3,
                                 "p99": float(np.percentile(host_ns, 99)) / 1e3, "max": float(host_ns.max()) / 1e3,
                                 "device_p50": float(np.percentile(dev_ns, 50)) / 1e3, "device_p99": float(np.percentile(dev_ns, 99)) / 1e3,
                                 "mode": "resident kernel, host-mapped mailbox, fused A1-A6+A7+A9+A11"}
    if rank == 0 and world == 1 and not args.no_cpu:
        r0, _, thr = cpu_reference_rate(32768, 1, 1)
        passes = int(max(1, min(64, round(r0 * 12 / rows))))          # ~12 s of CPU work over the same rows
        rate, dtc, thr = cpu_reference_rate(rows, passes, 0)
        # the "generous" CPU arm of SURVEY 8d: row-blocked forward (weights reused across rows), all cores, ~3 s
        rate_b, dtb, _ = cpu_reference_rate(rows, max(1, passes // 4), 0, blocked=True)
        # the control-loop step on one host core (C restatement of publish(), fp32 forward)
        from oracle import coracle
        cm1 = coracle.CModel(pkg.DEFAULT_MODEL)
        raws1 = [coracle.RawState.from_buffer_copy(bytes(r)) for r in synthetic_raw_states(capi, 512, seed=2)]
        coracle.closed_loop_latency_ns(cm1, raws1, 10_000)
        ns1 = coracle.closed_loop_latency_ns(cm1, raws1, 100_000)
        line["b1_cpu_port_us"] = {"steps": 100_000, "p50": float(np.percentile(ns1, 50)) / 1e3, "p99": float(np.percentile(ns1, 99)) / 1e3,
                                  "cores": 1, "kind": "port",
                                  "note": "C restatement of publish() (A1-A11, fp32 forward) on one host thread; ONNX Runtime itself is not installable here"}
        line["cpu_baseline_blocked"] = {"value": rate_b, "unit": UNIT, "cores": thr, "kind": "port",
                                        "sample": f"{max(1, passes // 4)} passes over {rows} rows in {dtb:.1f} s; row-blocked C forward "
                                                  "(oracle_mlp.c: orc_forward_blocked_f32), a batched CPU implementation the reference does not have"}
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
                                "sample": f"{passes} passes over {rows} rows of the same N(0,1) workload in {dtc:.1f} s; C restatement "
                                          "(oracle/oracle_mlp.c), batch-1 semantics per row, OpenMP over all host cores; "
                                          "ORT CPU EP itself is not installable here"}
    pb.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
