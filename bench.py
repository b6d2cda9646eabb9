#!/usr/bin/env python
"""Benchmark of the AMP hot path (BASELINE.json metric: "AMP obs samples/s (sample+obs+disc reward) at 1/2/4/8 B200;
% HBM peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input PER GPU (weak scaling, no data-path collective):

    collect_reference_motions(n)   fused history times -> frame/blend -> gather+lerp -> root slerp -> compute_obs -> (n, K*A)
    style reward over those rows   RunningStandardScaler -> MLP K*A-1024-512-1 on tcgen05 -> -log(max(1-sigmoid,1e-4))*2
    (+ env-step history update and the flat gradient all-reduce in the workloads that have them)

Workloads (``--workload``; shapes of the shipped clips, synthetic content, see humanoid_amp_b200/synthetic.py):

    refill_1m          BASELINE configs[3]: G1_walk shape, 1,000,000 samples x 2 history per GPU  (DEFAULT: the only
                       single-GPU config whose inputs exceed L2 and for which "% HBM peak" is meaningful)
    g1_walk_4096x2     BASELINE configs[1]: G1_walk, 4096 envs x K=2 (latency-bound: 2.9 MB per call; L2 flushed between steps)
    g1_dance_4096x10   BASELINE configs[2]: G1_dance (39 bodies), 4096 envs x K=10, reward over 16 rollouts x 4096 rows
    pooled_65536       BASELINE configs[4]: pooled humanoid walk+run+dance, 65536 envs sharded over the ranks (strong) +
                       NCCL all-reduce of the flat discriminator/policy/value gradient

Prints ONE JSON line (rank 0).  ``value`` = whole-job samples/s with inputs resident in HBM; ``e2e`` = the same through
the public Python API with HOST input buffers (H2D of times/ids and D2H of the rewards inside the timed region).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "amp_obs_samples_per_s(sample+obs+disc_reward)"
UNIT = "samples/s"

# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the DEFAULT workload
# (profiles/r01_ncu_full_v4_collect_cast.metrics.txt, profiles/r01_ncu_full_v5_step.metrics.txt); None for other workloads
NCU_TRAFFIC_BYTES = {
    "collect_reference_kernel": 16_283_392 + 605_972_736,  # 1 M samples x K=2 (algorithmic 680 MB; the tail still sat in L2)
    "disc_fused_kernel": 254_728_448 + 549_733_376,        # one 500 k-row launch: x_hat read + h1 slot write-backs
}

WORKLOADS = {
    # name: clip shape, samples per GPU, history K, reward rows multiplier, flush L2, extra stages
    "refill_1m": dict(clip="G1_walk", n=1_000_000, K=2, reward_mult=1, flush=False, env_step=False, allreduce=False, strong=False),
    "g1_walk_4096x2": dict(clip="G1_walk", n=4096, K=2, reward_mult=1, flush=True, env_step=True, allreduce=False, strong=False),
    "g1_dance_4096x10": dict(clip="G1_dance", n=4096, K=10, reward_mult=16, flush=True, env_step=True, allreduce=False, strong=False),
    "pooled_65536": dict(clip="pooled_humanoid", n=65536, K=2, reward_mult=1, flush=True, env_step=True, allreduce=True, strong=True),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------------------------
# clock sampling during the timed region (NVML in-process; nvidia-smi as a fallback)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
        0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }  # fmt: skip

    def __init__(self, device_index: int, period_s: float = 0.005):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self.period = period_s

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# synthetic workload construction
# ---------------------------------------------------------------------------------------------------------------------
def make_clip_files(tmp, clip, seed=0):
    from humanoid_amp_b200.synthetic import write_synthetic_clip

    if clip == "pooled_humanoid":
        return ",".join(write_synthetic_clip(os.path.join(tmp, f"humanoid_{n}.npz"), f"humanoid_{n}", seed + i) for i, n in enumerate(("walk", "run", "dance")))
    return write_synthetic_clip(os.path.join(tmp, f"{clip}.npz"), clip, seed)


def host_inputs(durations, n, seed):
    """ids uniform over the clips, times uniform in [0, duration) -- float64 / int64 host arrays."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, len(durations), n).astype(np.int64)
    times = rng.uniform(0.0, 1.0, n) * np.asarray(durations)[ids]
    return ids, times


def flops_per_row(in_features, h1=1024, h2=512):
    return 2.0 * (in_features * h1 + h1 * h2 + h2)


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU torch path
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(spec, clip_files, n_cpu, steps, warmup, seed=0):
    """Times sample + obs + style reward of the reference algorithm (oracle port, fp32 CPU torch, all host threads) on a
    bounded sample of the workload.  Returns (samples/s, ms/step, threads, description)."""
    from humanoid_amp_b200.robots import robot_for_clip
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params
    from oracle import OracleDiscriminator, OracleMotionLoader, env_oracle

    files = clip_files.split(",")
    ora = OracleMotionLoader(files)
    robot = robot_for_clip(ora.dof_names)
    K = spec["K"]
    width = K * robot.amp_observation_space
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
    disc = OracleDiscriminator(width, weights=W, biases=b)
    dof_idx = ora.get_dof_index(robot.joint_names)
    ref_idx = ora.get_body_index([robot.reference_body])[0]
    key_idx = ora.get_body_index(robot.key_body_names)
    ids, times = host_inputs(ora.durations, n_cpu, seed)
    disc.update_statistics(env_oracle.collect_reference_motions(ora, 2048, K, dof_idx, ref_idx, key_idx, current_times=times[:2048], motion_ids=ids[:2048]))

    def one():
        obs = env_oracle.collect_reference_motions(ora, n_cpu, K, dof_idx, ref_idx, key_idx, current_times=times, motion_ids=ids)
        rows = obs if spec["reward_mult"] == 1 else obs.repeat(spec["reward_mult"], 1)
        return disc.style_reward(rows)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_cpu / dt, dt * 1e3, torch.get_num_threads(), f"{n_cpu} of {spec['n']} samples x K={K} per step, oracle port (fp32 CPU torch), {steps} steps after {warmup} warm-up"


def torch_gpu_reference_pass(spec, clip_files, n, steps, warmup, dev, seed=0):
    """SURVEY.md 8d's optional second baseline: the SAME reference algorithm (oracle port) with its tensors on the B200, i.e.
    torch eager on the GPU -- what a user of the reference runs today.  Host index math + ~200 ATen launches + H2D copies per
    step, fp32 cuBLAS for the discriminator.  Timed with CUDA events around `steps` whole passes."""
    from humanoid_amp_b200.robots import robot_for_clip
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params
    from oracle import OracleDiscriminator, OracleMotionLoader, env_oracle

    ora = OracleMotionLoader(clip_files.split(","), device=dev)
    robot = robot_for_clip(ora.dof_names)
    K = spec["K"]
    width = K * robot.amp_observation_space
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
    disc = OracleDiscriminator(width, weights=W, biases=b, device=dev)
    dof_idx = ora.get_dof_index(robot.joint_names)
    ref_idx = ora.get_body_index([robot.reference_body])[0]
    key_idx = ora.get_body_index(robot.key_body_names)
    ids, times = host_inputs(ora.durations, n, seed)

    def one():
        obs = env_oracle.collect_reference_motions(ora, n, K, dof_idx, ref_idx, key_idx, current_times=times, motion_ids=ids)
        rows = obs if spec["reward_mult"] == 1 else obs.repeat(spec["reward_mult"], 1)
        return disc.style_reward(rows)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    return {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "kind": "oracle port on device=cuda (torch eager, fp32)",
            "sample": f"{n} samples x K={K} per step, {steps} steps after {warmup} warm-up, wall clock with synchronize on both sides"}


def run_reference(args, spec, rank, world):
    if rank != 0:
        return
    try:  # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as tmp:
        clip_files = make_clip_files(tmp, spec["clip"])
        n_cpu = min(spec["n"], 262144)  # bounded sample of the workload per step
        steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
        value, ms, threads, sample = cpu_reference_pass(spec, clip_files, n_cpu, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if spec["strong"] else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "clip_shape": spec["clip"], "K": spec["K"], "samples_per_step": n_cpu},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host": {"cpu_count": os.cpu_count(), "torch_threads": threads},
    }  # fmt: skip
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, spec, rank, world, local_rank):
    import torch.distributed as dist

    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params, synthetic_sim_state

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    peaks = measured_peaks()
    n = spec["n"] // world if spec["strong"] else spec["n"]
    if args.samples:
        n = args.samples
    K = spec["K"]
    tmp = tempfile.TemporaryDirectory()
    clip_files = make_clip_files(tmp.name, spec["clip"])
    loader = amp.MotionLoader(clip_files, dev)
    robot = amp.robot_for_clip(loader.dof_names)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=clip_files, num_envs=n if spec["env_step"] else 1, num_amp_observations=K, robot=robot), dev, motion_loader=loader)
    A = robot.amp_observation_space
    width = K * A
    reward_rows = n * spec["reward_mult"]

    # discriminator: torch.nn.Linear default init from seed 42, scaler statistics from reference observations
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
    ids_h, times_h = host_inputs(loader.durations, n, seed=1234 + rank)
    stats_obs = env.collect_reference_motions(4096, times_h[:4096], ids_h[:4096])
    mean = stats_obs.double().mean(dim=0)
    var = stats_obs.double().var(dim=0) + 1e-4
    disc = amp.AmpDiscriminator(width, device=dev, max_rows=reward_rows)
    disc.load(W, b, mean, var)

    # resident inputs / outputs
    times_d = torch.from_numpy(times_h).to(dev)
    ids_d = torch.from_numpy(ids_h).to(dev)
    obs = torch.empty((n, width), dtype=torch.float32, device=dev)
    rows_for_reward = obs if spec["reward_mult"] == 1 else torch.empty((reward_rows, width), dtype=torch.float32, device=dev).normal_()
    reward = torch.empty(reward_rows, dtype=torch.float32, device=dev)
    state = synthetic_sim_state(n, robot, dev, seed=7) if spec["env_step"] else None
    grads = bucket = None
    if spec["allreduce"]:
        # policy + value + discriminator of the humanoid config: one flat fp32 gradient bucket (SURVEY 8d cfg 5), averaged
        # over the ranks by the peer-memory kernel (csrc/amp_bucket.cu); AMP_B200_BENCH_NCCL=1 times NCCL instead
        n_disc = width * 1024 + 1024 + 1024 * 512 + 512 + 512 + 1
        if distributed and os.environ.get("AMP_B200_BENCH_NCCL") != "1":
            bucket = amp.GradientBucket(3 * n_disc, dev)
            grads = bucket.flat[: 3 * n_disc].normal_()
        else:
            grads = torch.randn(3 * n_disc, dtype=torch.float32, device=dev)

    def exchange_gradients():
        if grads is None or not distributed:
            return
        if bucket is not None:
            bucket.all_reduce_mean()
        else:
            dist.all_reduce(grads)
            grads.div_(world)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if spec["flush"] else None

    def step_resident():
        env.collect_reference_motions(n, times_d, ids_d, out=obs)
        if state is not None:
            env.update_amp_observations(*state)
        disc.style_reward(rows_for_reward, out=reward)
        exchange_gradients()

    launches_per_step = 1 + (1 if state is not None else 0) + disc.launch_count(reward_rows)
    launches_per_step += 1 if bucket is not None else 0  # the peer-memory all-reduce kernel

    def sync_all():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing --------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sync_all()

    def stage_obs():
        env.collect_reference_motions(n, times_d, ids_d, out=obs)
        if state is not None:
            env.update_amp_observations(*state)

    def stage_disc():
        disc.style_reward(rows_for_reward, out=reward)

    # Latency-bound workloads (a few MB per step): replay the two stages as CUDA graphs so the timed region measures the
    # GPU, not Python/ctypes launch overhead.  The big workload is launched eagerly.
    use_graph = flush_buf is not None and not args.no_graph
    g_obs = amp.capture_step(stage_obs, dev) if use_graph else None
    g_disc = amp.capture_step(stage_disc, dev) if use_graph else None
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
    whole0, whole1 = ev(), ev()
    with ClockSampler(local_rank) as clocks:
        sync_all()
        whole0.record()
        for s0, s1, s2 in marks:
            if flush_buf is not None:
                flush_buf.zero_()  # evict L2 between iterations (256 MB > 126 MB L2); outside the per-step events
            s0.record()
            g_obs.replay() if use_graph else stage_obs()
            s1.record()
            g_disc.replay() if use_graph else stage_disc()
            exchange_gradients()
            s2.record()
        whole1.record()
        sync_all()
    obs_ms = sum(a.elapsed_time(b) for a, b, _ in marks)
    disc_ms = sum(b.elapsed_time(c) for _, b, c in marks)
    if flush_buf is None:
        total_ms = whole0.elapsed_time(whole1)  # one bracket around exactly K steps
    else:
        total_ms = obs_ms + disc_ms  # the L2 flush between steps is excluded
    t = torch.tensor([total_ms, obs_ms, disc_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, obs_ms, disc_ms = (float(x) for x in t.tolist())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers -----------------------------------------------------------
    times_pin = torch.from_numpy(times_h).pin_memory()
    ids_pin = torch.from_numpy(ids_h).pin_memory()
    reward_host = torch.empty(reward_rows, dtype=torch.float32).pin_memory()

    def step_e2e():
        o = env.collect_reference_motions(n, times_pin, ids_pin, out=obs)  # H2D of times / ids inside
        if state is not None:
            env.update_amp_observations(*state)
        r = disc.style_reward(o if spec["reward_mult"] == 1 else rows_for_reward, out=reward)
        reward_host.copy_(r.view(-1), non_blocking=True)  # D2H of the step's result into pinned memory ...
        torch.cuda.current_stream(dev).synchronize()  # ... which the host then reads: one sync per step

    def run_serial(steps):
        for _ in range(steps):
            if flush_buf is not None:
                flush_buf.zero_()
            step_e2e()
        torch.cuda.synchronize(dev)

    # Pipelined variant (big workload only): the same per-step copies, but step i+1's inputs are prefetched and step i-1's
    # rewards are read back on copy streams while step i's kernels run (humanoid_amp_b200.pipeline).  EVERY step still copies
    # its own 16 B/sample in from pinned memory and its 4 B/sample result out, inside the timed region.
    pipelined = flush_buf is None
    if pipelined:
        pre = amp.InputPrefetcher(dev, n, depth=2)
        reader = amp.ResultReader(dev, depth=2)
        reward2 = [reward, torch.empty_like(reward)]

    def run_pipelined(steps):
        slot = pre.submit(times_pin, ids_pin)
        prev = None
        for i in range(steps):
            nxt = pre.submit(times_pin, ids_pin) if i + 1 < steps else None
            t_d, i_d = pre.acquire(slot)
            o = env.collect_reference_motions(n, t_d, i_d, out=obs)
            pre.release(slot)
            if state is not None:
                env.update_amp_observations(*state)
            r = disc.style_reward(o if spec["reward_mult"] == 1 else rows_for_reward, out=reward2[i & 1])
            ticket = reader.read_async(r.view(-1))
            if prev is not None:
                prev.wait()  # the host consumes step i-1's rewards while step i runs
            prev, slot = ticket, nxt
        prev.wait()
        torch.cuda.synchronize(dev)

    def time_e2e(run):
        run(max(2, args.warmup // 2))
        sync_all()
        t0 = time.perf_counter()
        run(args.steps)
        sec = time.perf_counter() - t0
        te = torch.tensor([sec], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_steps = args.steps
    e2e_s = time_e2e(run_serial)
    if flush_buf is not None:  # subtract the measured cost of the flushes themselves
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(e2e_steps):
            flush_buf.zero_()
        f1.record()
        torch.cuda.synchronize(dev)
        e2e_s = max(e2e_s - f0.elapsed_time(f1) * 1e-3, 1e-9)
    e2e_serial_value = n * world / (e2e_s / e2e_steps)
    e2e_value = n * world / (time_e2e(run_pipelined) / e2e_steps) if pipelined else e2e_serial_value

    # ---- roofline ------------------------------------------------------------------------------------------------------
    packed_bytes = loader.num_frames * ((A + 3) // 4 * 4) * 4
    obs_bytes = n * width * 4 + n * 16 + packed_bytes
    if state is not None:
        obs_bytes += n * ((2 * robot.num_joints + 25) * 4 + (K - 1) * A * 4 + K * A * 4)
    obs_gbs = obs_bytes / (obs_ms / args.steps * 1e-3) / 1e9
    disc_flops = reward_rows * flops_per_row(width)
    disc_tflops = disc_flops / (disc_ms / args.steps * 1e-3) / 1e12
    tensor_peak = peaks["bf16_sustained"]
    roofline = {
        "kernel": "disc_fused_kernel (tcgen05 two-layer discriminator + style reward) + normalise_cast_kernel: 2 launches per chunk", "bound": "tensor",
        "achieved": disc_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": disc_tflops / tensor_peak,
        "traffic": NCU_TRAFFIC_BYTES["disc_fused_kernel"] if args.workload == "refill_1m" and not args.samples else None,
        "traffic_note": "DRAM bytes of ONE fused launch (500 k rows; two per step), ncu --set full",
        "algorithmic_flops_per_launch_group": disc_flops, "ms": disc_ms / args.steps, "peak_source": peaks["source"] + ", sustained bf16",
    }
    roofline_hbm = {
        "kernel": "collect_reference_kernel (+ obs_step_kernel)" if state is not None else "collect_reference_kernel", "bound": "hbm",
        "achieved": obs_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": obs_gbs / peaks["hbm_gbs"],
        "traffic": NCU_TRAFFIC_BYTES["collect_reference_kernel"] if args.workload == "refill_1m" and not args.samples else None,
        "algorithmic_bytes_per_launch": obs_bytes, "ms": obs_ms / args.steps, "us_per_call": obs_ms / args.steps * 1e3, "peak_source": peaks["source"],
    }

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = min(n, 262144)  # bounded sample: ~0.6 s per step on 16 cores -> ~6-10 s of CPU work in total
            v, ms, threads, sample = cpu_reference_pass(spec, clip_files, n_cpu, steps=10 if n_cpu >= 65536 else 30, warmup=1)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "ms_per_step": ms}
        torch_gpu = None
        if world == 1 and args.torch_gpu_baseline:
            del obs, reward  # give the eager pass (tens of full-size temporaries) the memory back
            torch.cuda.empty_cache()
            torch_gpu = torch_gpu_reference_pass(spec, clip_files, n, steps=5, warmup=2, dev=dev)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if spec["strong"] else "weak", "vs_baseline": None,
            "dtype": "f32 (sample/obs, f64 index math) + bf16 tensor-core MLP with f32 accumulate", "data": "synthetic",
            "config": {
                "workload": args.workload, "clip_shape": spec["clip"], "frames": loader.num_frames, "samples_per_gpu": n, "K": K,
                "amp_obs_width": A, "disc": f"{width}-1024-512-1", "reward_rows_per_step": reward_rows,
                "l2": "inputs+outputs larger than L2 (no flush)" if flush_buf is None else "256 MB write between timed steps flushes L2",
                "launch": "cuda_graph replay per stage" if use_graph else "eager (Python -> ctypes -> C ABI)",
            },
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
            **({"torch_gpu_baseline": torch_gpu} if torch_gpu is not None else {}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": reward_rows * 4,
                    "mode": "double-buffered host staging: H2D of step i+1 and D2H of step i-1 overlap step i's kernels" if pipelined
                    else "serial: H2D -> kernels -> D2H -> sync every step",
                    "serial_value": e2e_serial_value},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(),
            "stage_ms": {"sample+obs": obs_ms / args.steps, "disc_reward": disc_ms / args.steps},
        }  # fmt: skip
        print(json.dumps(line), file=JSON_OUT, flush=True)
    tmp.cleanup()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


JSON_OUT = sys.stdout  # the ONE JSON line goes here; everything else printed during a run (the loaders' two
                       # "Loading ..." lines, like the reference's MotionLoader) is sent to stderr, see main()


def main():
    global JSON_OUT
    JSON_OUT, sys.stdout = sys.stdout, sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="refill_1m")
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-gpu-baseline", action="store_true",
                    help="also time the reference algorithm as torch eager on this GPU (SURVEY 8d's optional second baseline)")
    ap.add_argument("--no-graph", action="store_true", help="launch the small workloads eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, spec, rank, world)
        return
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    run_ours(args, spec, rank, world, local_rank)


if __name__ == "__main__":
    main()
