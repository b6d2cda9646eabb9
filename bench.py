#!/usr/bin/env python
"""Benchmark of the AMP hot path (BASELINE.json metric: "AMP obs samples/s (sample+obs+disc reward) at 1/2/4/8 B200;
% HBM peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input PER GPU:

    collect_reference_motions(n)   fused history times -> frame/blend -> gather+lerp -> root slerp -> compute_obs -> (n, K*A)
    style reward over those rows   RunningStandardScaler -> MLP K*A-1024-512-1 on tcgen05 -> -log(max(1-sigmoid,1e-4))*2
    (+ the env-step observation update in the workloads that have envs)
    (+ at WORLD_SIZE > 1: the gradient exchange the reference runs per mini-batch, train.py:184-196 -- one flat fp32 bucket
       of policy + value + discriminator gradients averaged over the ranks by the peer-memory kernel of csrc/amp_bucket.cu)

Workloads (``--workload``; shapes of the shipped clips, synthetic content, see humanoid_amp_b200/synthetic.py):

    refill_1m          BASELINE configs[3]: G1_walk shape, 1,000,000 samples x 2 history per GPU  (DEFAULT: the only
                       single-GPU config whose inputs exceed L2 and for which "% HBM peak" is meaningful)
    g1_walk_4096x2     BASELINE configs[1]: G1_walk, 4096 envs x K=2 (latency-bound: 2.9 MB per call; L2 flushed between steps)
    g1_dance_4096x10   BASELINE configs[2]: G1_dance (39 bodies), 4096 envs x K=10, reward over 16 rollouts x 4096 rows
    humanoid_walk_4096x2  BASELINE configs[0]: humanoid_walk shape, 4096 samples x K=2 (the reference's CPU-runnable case)
    pooled_65536       BASELINE configs[4]: pooled humanoid walk+run+dance, 65536 envs sharded over the ranks (strong) +
                       the gradient exchange

Prints ONE JSON line (rank 0).  ``value`` = whole-job samples/s with inputs resident in HBM; ``e2e`` = the same through
the public Python API with HOST input buffers (H2D of times/ids and D2H of the rewards inside the timed region).  The
default run also measures the latency-bound BASELINE configs and reports them under ``other_workloads``.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "amp_obs_samples_per_s(sample+obs+disc_reward)"
UNIT = "samples/s"

WORKLOADS = {
    # name: clip shape, samples per GPU, history K, reward rows multiplier, flush L2, env step, strong scaling
    "refill_1m": dict(clip="G1_walk", n=1_000_000, K=2, reward_mult=1, flush=False, env_step=False, strong=False),
    "g1_walk_4096x2": dict(clip="G1_walk", n=4096, K=2, reward_mult=1, flush=True, env_step=True, strong=False),
    "g1_dance_4096x10": dict(clip="G1_dance", n=4096, K=10, reward_mult=16, flush=True, env_step=True, strong=False),
    "humanoid_walk_4096x2": dict(clip="humanoid_walk", n=4096, K=2, reward_mult=1, flush=True, env_step=False, strong=False),
    "pooled_65536": dict(clip="pooled_humanoid", n=65536, K=2, reward_mult=1, flush=True, env_step=True, strong=True),
}
# policy + value + discriminator parameters of the G1-dance config: the bucket the reference all-reduces
EXCHANGE_FLOATS = 645_000 + 631_000 + (830 * 1024 + 1024 + 1024 * 512 + 512 + 512 + 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel: str, workload: str):
    """DRAM bytes per launch of ``kernel`` from the committed ``ncu --set full`` capture of this workload
    (``profiles/traffic.json``: bytes, capture file and the commit it was taken at), or None when there is none: the
    figure is NOT measured in this run and is reported with its provenance instead of as a constant in the code."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        table = json.load(f)
    rec = table.get(workload, {}).get(kernel)
    return (rec["dram_bytes_per_launch"], rec) if rec else (None, None)


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# clock sampling during the timed region (NVML in-process)
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


def workload_config(name, spec, n, frames, A):
    """The workload as both arms report it (the driver compares the two ``config`` objects)."""
    width = spec["K"] * A
    return {
        "workload": name, "clip_shape": spec["clip"], "frames": int(frames), "samples_per_gpu": int(n), "K": spec["K"], "amp_obs_width": A,
        "disc": f"{width}-1024-512-1", "reward_rows_per_step": int(n * spec["reward_mult"]),
        "l2": "256 MB write between timed steps flushes L2" if spec["flush"] else "inputs+outputs larger than L2 (no flush)",
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------------------------------------------------
def make_cpu_reference(spec, clip_files, device="cpu"):
    """Returns (one_step(times, ids) -> rewards, kind, description, durations, frames, A).

    kind "reference": ``oracle/_ref`` is present (installed by the committed recipe ``oracle/build_ref.py`` from
    ``/root/reference``; travels to the GPU box with the snapshot) -- the UNMODIFIED reference ``MotionLoader``
    (``motions/motion_loader.py:87-430``) and the reference's own ``collect_reference_motions`` / ``compute_obs`` text
    (``g1_amp_env.py:445-497, 535-561``) run on CPU torch.  The discriminator stage is skrl's expression restated
    (``oracle/disc_oracle.py``): skrl is a third-party dependency that is neither vendored by the reference nor installed.
    kind "port": ``oracle/_ref`` is absent, the whole path is the oracle restatement.
    """
    from humanoid_amp_b200.robots import robot_for_clip
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params
    from oracle import OracleDiscriminator, OracleMotionLoader, env_oracle, ref_harness

    K = spec["K"]
    if ref_harness.available():
        loader = ref_harness.reference_motion_loader_class()(clip_files, device)
        robot = robot_for_clip(loader.dof_names)
        env = ref_harness.make_ref_env(loader, robot, 1, K)
        kind = "reference"
        what = "unmodified reference MotionLoader + collect_reference_motions/compute_obs text (oracle/_ref), restated skrl discriminator"

        def collect(times, ids):
            return env.collect_reference_motions(len(times), times, ids)
    else:
        loader = OracleMotionLoader(clip_files.split(","), device=device)
        robot = robot_for_clip(loader.dof_names)
        dof_idx = loader.get_dof_index(robot.joint_names)
        ref_idx = loader.get_body_index([robot.reference_body])[0]
        key_idx = loader.get_body_index(robot.key_body_names)
        kind = "port"
        what = "oracle port (oracle/_ref absent)"

        def collect(times, ids):
            return env_oracle.collect_reference_motions(loader, len(times), K, dof_idx, ref_idx, key_idx, current_times=times, motion_ids=ids)

    A = robot.amp_observation_space
    width = K * A
    W, b = skrl_style_discriminator_params(width, seed=42, logit_gain=5.0)
    disc = OracleDiscriminator(width, weights=W, biases=b, device=device)
    ids0, times0 = host_inputs(loader.durations, 2048, 99)
    disc.update_statistics(collect(times0, ids0))

    def one(times, ids):
        obs = collect(times, ids)
        rows = obs if spec["reward_mult"] == 1 else obs.repeat(spec["reward_mult"], 1)
        return disc.style_reward(rows)

    return one, kind, what, np.asarray(loader.durations), int(loader.num_frames), A


def cpu_reference_pass(spec, clip_files, n_cpu, steps, warmup, seed=0):
    """Times sample + obs + style reward of the reference's CPU path (fp32 CPU torch, all host threads) on ``n_cpu`` samples
    per step.  Returns a dict(value, ms_per_step, cores, kind, sample, frames, A)."""
    one, kind, what, durations, frames, A = make_cpu_reference(spec, clip_files)
    ids, times = host_inputs(durations, n_cpu, seed)
    for _ in range(warmup):
        one(times, ids)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(times, ids)
    dt = (time.perf_counter() - t0) / steps
    sample = f"{n_cpu} of {spec['n']} samples x K={spec['K']} per step, {what}, fp32 CPU torch, {steps} steps after {warmup} warm-up"
    return dict(value=n_cpu / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(), kind=kind, sample=sample, frames=frames, A=A)


def torch_gpu_reference_pass(spec, clip_files, n, steps, warmup, dev, seed=0):
    """SURVEY.md 8d's second baseline: the SAME reference code with its tensors on the B200, i.e. torch eager on the GPU --
    what a user of the reference runs today.  Host index math + ~200 ATen launches + H2D copies per step, fp32 cuBLAS for the
    discriminator.  Wall clock around ``steps`` whole passes with a synchronize on both sides."""
    one, kind, what, durations, _frames, _A = make_cpu_reference(spec, clip_files, device=dev)
    ids, times = host_inputs(durations, n, seed)
    for _ in range(warmup):
        one(times, ids)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(times, ids)
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    return {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "kind": f"{kind} code on device=cuda (torch eager, fp32)",
            "sample": f"{n} samples x K={spec['K']} per step, {what}, {steps} steps after {warmup} warm-up"}


def set_all_host_threads():
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)


def run_reference(args, name, spec, rank, world):
    """``--impl reference``: the reference's CPU implementation on the SAME workload (n, K, steps, warm-up) as the GPU arm.
    A step is only bounded (fewer samples) when the whole run would otherwise exceed ~4 minutes; the line says so."""
    if rank != 0:
        return
    set_all_host_threads()
    with tempfile.TemporaryDirectory() as tmp:
        clip_files = make_clip_files(tmp, spec["clip"])
        n = args.samples or spec["n"]
        probe = cpu_reference_pass(spec, clip_files, min(n, 32768), steps=1, warmup=1)
        budget_s = 240.0
        n_cpu = n
        est = n / probe["value"] * (args.steps + args.warmup)
        if est > budget_s:
            n_cpu = max(4096, int(n * budget_s / est) // 4096 * 4096)
        r = cpu_reference_pass(spec, clip_files, n_cpu, args.steps, args.warmup)
    cfg = workload_config(name, spec, n, r["frames"], r["A"])
    if n_cpu != n:
        cfg["bounded_samples_per_step"] = n_cpu
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if spec["strong"] else "weak", "vs_baseline": None,
        "dtype": "f32 (f64 index math)", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host": {"cpu_count": os.cpu_count(), "torch_threads": r["cores"]},
    }  # fmt: skip
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def ev():
    return torch.cuda.Event(enable_timing=True)


def build_workload(name, spec, args, rank, world, dev):
    """Everything one workload needs, resident on the device: loader, env, discriminator, inputs, outputs."""
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params, synthetic_sim_state

    w = SimpleNamespace(name=name, spec=spec, dev=dev, amp=amp)
    n = spec["n"] // world if spec["strong"] else spec["n"]
    if args.samples and name == args.workload:
        n = args.samples
    w.n, w.K = n, spec["K"]
    w.tmp = tempfile.TemporaryDirectory()
    w.clip_files = make_clip_files(w.tmp.name, spec["clip"])
    w.loader = amp.MotionLoader(w.clip_files, dev)
    w.robot = amp.robot_for_clip(w.loader.dof_names)
    w.env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=w.clip_files, num_envs=n if spec["env_step"] else 1, num_amp_observations=w.K, robot=w.robot), dev, motion_loader=w.loader)
    w.A = w.robot.amp_observation_space
    w.width = w.K * w.A
    w.reward_rows = n * spec["reward_mult"]
    # discriminator: torch.nn.Linear default init from seed 42, scaler statistics from reference observations
    W, b = skrl_style_discriminator_params(w.width, seed=42, logit_gain=5.0)
    w.ids_h, w.times_h = host_inputs(w.loader.durations, n, seed=1234 + rank)
    stats_obs = w.env.collect_reference_motions(min(n, 4096), w.times_h[:4096], w.ids_h[:4096])
    w.disc = amp.AmpDiscriminator(w.width, device=dev, max_rows=w.reward_rows)
    w.disc.load(W, b, stats_obs.double().mean(dim=0), stats_obs.double().var(dim=0) + 1e-4)
    # resident inputs / outputs
    w.times_d = torch.from_numpy(w.times_h).to(dev)
    w.ids_d = torch.from_numpy(w.ids_h).to(dev)
    w.obs = torch.empty((n, w.width), dtype=torch.float32, device=dev)
    w.rows_for_reward = w.obs if spec["reward_mult"] == 1 else torch.empty((w.reward_rows, w.width), dtype=torch.float32, device=dev).normal_()
    w.reward = torch.empty(w.reward_rows, dtype=torch.float32, device=dev)
    w.state = synthetic_sim_state(n, w.robot, dev, seed=7) if spec["env_step"] else None
    w.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if spec["flush"] else None
    w.launches_per_step = 1 + (1 if w.state is not None else 0) + w.disc.launch_count(w.reward_rows)
    return w


def stage_obs(w):
    w.env.collect_reference_motions(w.n, w.times_d, w.ids_d, out=w.obs)
    if w.state is not None:
        w.env.update_amp_observations(*w.state)


def stage_disc(w):
    w.disc.style_reward(w.rows_for_reward, out=w.reward)


def time_resident(w, steps, warmup, sync_all, exchange=None, use_graph=True):
    """Device-resident timing with CUDA events on the launching stream.  Returns (total_ms, obs_ms, disc_ms, exchange_ms)
    summed over ``steps``.  Latency-bound workloads (L2 flushed between steps) replay the WHOLE step as one CUDA graph for
    the total and, in a second pass, one graph per stage for the breakdown; the big workload is launched eagerly."""
    amp = w.amp

    def whole():
        stage_obs(w)
        stage_disc(w)

    for _ in range(warmup):
        whole()
        if exchange:
            exchange()
    sync_all()
    graphs = w.flush_buf is not None and use_graph
    g_whole = amp.capture_step(whole, w.dev) if graphs else None
    g_obs = amp.capture_step(lambda: stage_obs(w), w.dev) if graphs else None
    g_disc = amp.capture_step(lambda: stage_disc(w), w.dev) if graphs else None
    marks = [(ev(), ev(), ev(), ev()) for _ in range(steps)]
    whole0, whole1 = ev(), ev()
    sync_all()
    whole0.record()
    for s0, s1, s2, s3 in marks:
        if w.flush_buf is not None:
            w.flush_buf.zero_()  # evict L2 between iterations (256 MB > 126 MB L2); outside the per-step events
        s0.record()
        if graphs:
            g_whole.replay()
            s1.record()
        else:
            stage_obs(w)
            s1.record()
            stage_disc(w)
        s2.record()
        if exchange:
            exchange()
        s3.record()
    whole1.record()
    sync_all()
    ex_ms = sum(c.elapsed_time(d) for _, _, c, d in marks)
    if w.flush_buf is None:
        total_ms = whole0.elapsed_time(whole1)  # one bracket around exactly K steps
        obs_ms = sum(a.elapsed_time(b) for a, b, _, _ in marks)
        disc_ms = sum(b.elapsed_time(c) for _, b, c, _ in marks)
    else:
        total_ms = sum(a.elapsed_time(d) for a, _, _, d in marks)  # the L2 flush between steps is excluded
        obs_ms = disc_ms = 0.0
        if graphs:  # second pass: the same step as two graphs, for the per-stage breakdown only
            parts = [(ev(), ev(), ev()) for _ in range(steps)]
            for p0, p1, p2 in parts:
                w.flush_buf.zero_()
                p0.record()
                g_obs.replay()
                p1.record()
                g_disc.replay()
                p2.record()
            torch.cuda.synchronize(w.dev)
            obs_ms = sum(a.elapsed_time(b) for a, b, _ in parts)
            disc_ms = sum(b.elapsed_time(c) for _, b, c in parts)
    return total_ms, obs_ms, disc_ms, ex_ms


def time_e2e(w, steps, warmup, sync_all):
    """The same metric end to end through the public Python API with HOST buffers.  Every step copies its own 16 B/sample of
    inputs from pinned host memory and its 4 B/sample of rewards back, inside the timed region.  Returns (value_seconds
    per step, mode, serial_seconds per step)."""
    amp, dev, n = w.amp, w.dev, w.n
    times_pin = torch.from_numpy(w.times_h).pin_memory()
    ids_pin = torch.from_numpy(w.ids_h).pin_memory()
    reward_host = torch.empty(w.reward_rows, dtype=torch.float32).pin_memory()
    reward_src = w.obs if w.spec["reward_mult"] == 1 else w.rows_for_reward

    def step_serial():
        o = w.env.collect_reference_motions(n, times_pin, ids_pin, out=w.obs)  # H2D of times / ids inside
        if w.state is not None:
            w.env.update_amp_observations(*w.state)
        r = w.disc.style_reward(o if w.spec["reward_mult"] == 1 else reward_src, out=w.reward)
        reward_host.copy_(r.view(-1), non_blocking=True)  # D2H of the step's result into pinned memory ...
        torch.cuda.current_stream(dev).synchronize()  # ... which the host then reads: one sync per step

    def run_serial(k):
        for _ in range(k):
            if w.flush_buf is not None:
                w.flush_buf.zero_()
            step_serial()
        torch.cuda.synchronize(dev)

    def clock(run, k):
        run(max(2, warmup // 2))
        sync_all()
        t0 = time.perf_counter()
        run(k)
        return time.perf_counter() - t0

    def minus_flush(sec, k):
        if w.flush_buf is None:
            return sec
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(k):
            w.flush_buf.zero_()
        f1.record()
        torch.cuda.synchronize(dev)
        return max(sec - f0.elapsed_time(f1) * 1e-3, 1e-9)

    serial = minus_flush(clock(run_serial, steps), steps) / steps
    if w.flush_buf is None:
        # big workload: double-buffered host staging (humanoid_amp_b200.pipeline): H2D of step i+1 and D2H of step i-1 run on
        # copy streams under step i's kernels
        pre = amp.InputPrefetcher(dev, n, depth=2)
        reader = amp.ResultReader(dev, depth=2)
        reward2 = [w.reward, torch.empty_like(w.reward)]

        def run_pipelined(k):
            slot = pre.submit(times_pin, ids_pin)
            prev = None
            for i in range(k):
                nxt = pre.submit(times_pin, ids_pin) if i + 1 < k else None
                t_d, i_d = pre.acquire(slot)
                o = w.env.collect_reference_motions(n, t_d, i_d, out=w.obs)
                pre.release(slot)
                r = w.disc.style_reward(o, out=reward2[i & 1])
                ticket = reader.read_async(r.view(-1))
                if prev is not None:
                    prev.wait()  # the host consumes step i-1's rewards while step i runs
                prev, slot = ticket, nxt
            prev.wait()
            torch.cuda.synchronize(dev)

        return clock(run_pipelined, steps) / steps, "double-buffered host staging: H2D of step i+1 and D2H of step i-1 overlap step i's kernels", serial

    # latency-bound workloads: the WHOLE step -- H2D of the pinned inputs, every kernel, D2H of the rewards -- is ONE CUDA
    # graph; per step the host writes its inputs into the pinned buffers, launches the graph and synchronises once
    times_dev, ids_dev = torch.empty_like(w.times_d), torch.empty_like(w.ids_d)

    def body():
        times_dev.copy_(times_pin, non_blocking=True)
        ids_dev.copy_(ids_pin, non_blocking=True)
        w.env.collect_reference_motions(n, times_dev, ids_dev, out=w.obs)
        if w.state is not None:
            w.env.update_amp_observations(*w.state)
        r = w.disc.style_reward(reward_src, out=w.reward)
        reward_host.copy_(r.view(-1), non_blocking=True)

    graph = amp.capture_step(body, dev)
    stream = torch.cuda.current_stream(dev)

    def run_graph(k):
        for _ in range(k):
            w.flush_buf.zero_()
            graph.replay()
            stream.synchronize()

    return minus_flush(clock(run_graph, steps), steps) / steps, "one CUDA graph per step (H2D of pinned inputs + kernels + D2H of rewards), one host sync per step", serial


def roofline_records(w, steps, obs_ms, disc_ms, total_timed_s, peaks):
    n, K, A = w.n, w.K, w.A
    packed_bytes = w.loader.num_frames * ((A + 3) // 4 * 4) * 4
    obs_bytes = n * w.width * 4 + n * 16 + packed_bytes
    if w.state is not None:
        obs_bytes += n * ((2 * w.robot.num_joints + 25) * 4 + (K - 1) * A * 4 + K * A * 4)
    obs_s = max(obs_ms / steps * 1e-3, 1e-12)
    disc_s = max(disc_ms / steps * 1e-3, 1e-12)
    obs_gbs = obs_bytes / obs_s / 1e9
    disc_flops = w.reward_rows * flops_per_row(w.width)
    disc_tflops = disc_flops / disc_s / 1e12
    # a timed region shorter than ~1 s runs at burst clocks (no power throttling yet): the burst cuBLAS figure is the matching
    # denominator; longer regions are compared with the sustained one.  Both fractions are printed.
    sustained = total_timed_s >= 1.0
    tensor_peak = peaks["bf16_sustained"] if sustained else peaks["bf16_burst"]
    traffic, traffic_rec = ncu_traffic("disc_fused_kernel", w.name)
    roofline = {
        "kernel": "disc_fused_kernel (tcgen05: scaler + bf16 cast + layer 1 + layer 2 + w3 dot + style reward in one launch)", "bound": "tensor",
        "achieved": disc_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": disc_tflops / tensor_peak,
        "frac_of_burst": disc_tflops / peaks["bf16_burst"], "frac_of_sustained": disc_tflops / peaks["bf16_sustained"],
        "peak_kind": "sustained (timed region >= 1 s)" if sustained else "burst (timed region < 1 s)",
        "traffic": traffic, "traffic_source": traffic_rec, "algorithmic_flops_per_step": disc_flops, "launches_per_step": w.disc.launch_count(w.reward_rows),
        "ms": disc_ms / steps, "peak_source": peaks["source"],
    }  # fmt: skip
    traffic_h, traffic_h_rec = ncu_traffic("collect_reference_kernel", w.name)
    roofline_hbm = {
        "kernel": "collect_reference_kernel (+ obs_step_kernel)" if w.state is not None else "collect_reference_kernel", "bound": "hbm",
        "achieved": obs_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": obs_gbs / peaks["hbm_gbs"],
        "traffic": traffic_h, "traffic_source": traffic_h_rec, "algorithmic_bytes_per_launch": obs_bytes, "ms": obs_ms / steps,
        "us_per_call": obs_ms / steps * 1e3, "peak_source": peaks["source"],
    }  # fmt: skip
    return roofline, roofline_hbm


def measure_secondary(name, args, rank, world, dev, sync_all, steps=40):
    """A compact record of one of the other BASELINE configs (latency-bound: whole step as one CUDA graph)."""
    spec = WORKLOADS[name]
    w = build_workload(name, spec, args, rank, world, dev)
    total_ms, obs_ms, disc_ms, _ = time_resident(w, steps, 5, sync_all, use_graph=not args.no_graph)
    t = torch.tensor([total_ms, obs_ms, disc_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, obs_ms, disc_ms = (float(x) for x in t.tolist())
    e2e_s, mode, serial_s = time_e2e(w, steps, 5, sync_all)
    te = torch.tensor([e2e_s, serial_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, serial_s = (float(x) for x in te.tolist())
    peaks = measured_peaks()
    roof, roof_h = roofline_records(w, steps, obs_ms, disc_ms, total_ms * 1e-3, peaks)
    rec = {
        "config": workload_config(name, spec, w.n, w.loader.num_frames, w.A), "steps": steps, "us_per_step": total_ms / steps * 1e3,
        "value": w.n * world / (total_ms / steps * 1e-3), "unit": UNIT, "scaling": "strong" if spec["strong"] else "weak",
        "e2e": {"value": w.n * world / e2e_s, "us_per_step": e2e_s * 1e6, "mode": mode, "serial_us_per_step": serial_s * 1e6,
                "h2d_bytes_per_step": w.n * 16, "d2h_bytes_per_step": w.reward_rows * 4},
        "stage_us": {"sample+obs": obs_ms / steps * 1e3, "disc_reward": disc_ms / steps * 1e3},
        "disc_frac_of_burst": roof["frac_of_burst"], "disc_tflops": roof["achieved"], "obs_gbs": roof_h["achieved"],
        "launch": "whole step replayed as one CUDA graph" if not args.no_graph else "eager",
        "gpu_launches_per_step": w.launches_per_step,
    }  # fmt: skip
    w.tmp.cleanup()
    del w
    torch.cuda.empty_cache()
    return rec


def run_ours(args, name, spec, rank, world, local_rank):
    import torch.distributed as dist

    import humanoid_amp_b200 as amp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    def sync_all():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peaks = measured_peaks()
    w = build_workload(name, spec, args, rank, world, dev)
    n, K, A = w.n, w.K, w.A

    # ---- the gradient exchange (world > 1): the reference all-reduces policy + value + discriminator gradients per mini-batch
    # (train.py:184-196 -> skrl Model.reduce_parameters); here ONE flat bucket and one peer-memory kernel per rank ----------
    bucket = exchange = collective = None
    if distributed and not args.no_exchange:
        bucket = amp.GradientBucket(EXCHANGE_FLOATS, dev)
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        src = torch.randn(EXCHANGE_FLOATS, device=dev, generator=g) * (1.0 + rank)
        bucket.flat[:EXCHANGE_FLOATS].copy_(src)
        want = src.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        want /= world
        bucket.all_reduce_mean()  # checked ONCE against NCCL, outside the timed region
        torch.cuda.synchronize(dev)
        diff = torch.tensor([float((bucket.flat[:EXCHANGE_FLOATS] - want).abs().max())], dtype=torch.float64, device=dev)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        ref0 = bucket.flat[:EXCHANGE_FLOATS].clone()
        dist.broadcast(ref0, 0)
        same = torch.tensor([1 if torch.equal(ref0, bucket.flat[:EXCHANGE_FLOATS]) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        kernel = ("allreduce_mean_switch_kernel (csrc/amp_bucket.cu: multimem.ld_reduce / multimem.st on an NVSwitch multicast mapping of the "
                  "ranks' buckets, in place)" if bucket.in_switch else
                  "allreduce_mean_bulk_kernel (csrc/amp_bucket.cu: two-shot, cp.async.bulk over NVLink peer memory, in place)")
        collective = {"kernel": kernel, "in_switch": bucket.in_switch,
                      "replaces": "skrl Model.reduce_parameters: NCCL all_reduce(SUM) + divide (reference train.py:184-196)",
                      "floats": EXCHANGE_FLOATS, "max_abs_diff_vs_nccl": float(diff.item()), "bitwise_identical_on_all_ranks": bool(same.item())}  # fmt: skip
        bucket.flat[:EXCHANGE_FLOATS].copy_(src)
        exchange = bucket.all_reduce_mean
        w.launches_per_step += 1

    # ---- device-resident timing ------------------------------------------------------------------------------------------
    with ClockSampler(local_rank) as clocks:
        total_ms, obs_ms, disc_ms, ex_ms = time_resident(w, args.steps, args.warmup, sync_all, exchange=exchange, use_graph=not args.no_graph)
    t = torch.tensor([total_ms, obs_ms, disc_ms, ex_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, obs_ms, disc_ms, ex_ms = (float(x) for x in t.tolist())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    if collective is not None:
        scratch = torch.randn(EXCHANGE_FLOATS, device=dev)

        def nccl():
            dist.all_reduce(scratch, op=dist.ReduceOp.SUM)
            scratch.div_(world)

        def timed(fn, iters=30):
            for _ in range(5):
                fn()
            sync_all()
            a, b = ev(), ev()
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize(dev)
            x = torch.tensor([a.elapsed_time(b) / iters * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            return float(x.item())

        collective["us_in_step"] = ex_ms / args.steps * 1e3  # behind the discriminator kernel, ranks skewed by their own steps
        collective["us"] = timed(bucket.all_reduce_mean)  # back to back, ranks in lock step
        collective["nccl_us"] = timed(nccl)
        st = torch.tensor([bucket.poll_status()], device=dev)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        collective["status"] = int(st.item())

    # ---- end to end through the public API with host buffers ---------------------------------------------------------------
    e2e_s, e2e_mode, serial_s = time_e2e(w, args.steps, args.warmup, sync_all)
    te = torch.tensor([e2e_s, serial_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, serial_s = (float(x) for x in te.tolist())

    roofline, roofline_hbm = roofline_records(w, args.steps, obs_ms, disc_ms, total_ms * 1e-3, peaks)
    cfg = workload_config(name, spec, n, w.loader.num_frames, A)
    launches = w.launches_per_step
    frames = w.loader.num_frames
    clip_files = w.clip_files

    # ---- the other BASELINE configs (compact records), the CPU baseline and the torch-eager-on-GPU baseline --------------------
    others = {}
    if not args.no_others and name == "refill_1m":
        del w.obs, w.reward, w.rows_for_reward  # free the 1 M-row buffers
        torch.cuda.empty_cache()
        names = ["pooled_65536"] if distributed else ["g1_walk_4096x2", "g1_dance_4096x10", "humanoid_walk_4096x2"]
        for other in names:
            others[other] = measure_secondary(other, args, rank, world, dev, sync_all)

    if rank == 0:
        cpu = torch_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            set_all_host_threads()
            n_cpu = min(n, 262144)  # bounded sample: ~0.6 s per step on 16 cores -> ~6-10 s of CPU work in total
            r = cpu_reference_pass(spec, clip_files, n_cpu, steps=10 if n_cpu >= 65536 else 30, warmup=1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
        if world == 1 and not args.no_torch_gpu_baseline:
            torch.cuda.empty_cache()
            try:
                torch_gpu = torch_gpu_reference_pass(spec, clip_files, n, steps=5, warmup=2, dev=dev)
                torch_gpu["speedup_of_this_path"] = value / torch_gpu["value"]
            except Exception as exc:  # the eager pass holds tens of full-size temporaries: report instead of failing the line
                torch_gpu = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if spec["strong"] else "weak", "vs_baseline": None,
            "dtype": "f32 (sample/obs, f64 index math) + bf16 tensor-core MLP with f32 accumulate", "data": "synthetic",
            "config": cfg, "launch_mode": "whole step replayed as one CUDA graph" if (spec["flush"] and not args.no_graph) else "eager (Python -> ctypes -> C ABI)",
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu,
            "e2e": {"value": n * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": w.reward_rows * 4,
                    "mode": e2e_mode, "serial_value": n * world / serial_s},
            "gpu_launches": launches * args.steps, "clocks": clocks.summary(),
            "stage_ms": {"sample+obs": obs_ms / args.steps, "disc_reward": disc_ms / args.steps, "gradient_exchange": ex_ms / args.steps},
            "other_workloads": others, "commit": git_head(),
        }  # fmt: skip
        if collective is not None:
            line["collective"] = collective
            line["value_without_exchange"] = n * world / ((total_ms - ex_ms) / args.steps * 1e-3)
        print(json.dumps(line), file=JSON_OUT, flush=True)
    w.tmp.cleanup()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


JSON_OUT = sys.stdout  # the ONE JSON line goes here; everything else printed during a run (the loaders' two
                       # "Loading ..." lines, like the reference's MotionLoader) is sent to stderr, see main()


def main():
    global JSON_OUT
    # stdout carries ONLY the JSON line: Python-level prints go to stderr, and so does everything native code writes to file
    # descriptor 1 (NCCL prints its version banner there) -- the line itself is written to a duplicate of the original fd
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="refill_1m")
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true", help="skip timing the reference code as torch eager on this GPU")
    ap.add_argument("--no-others", action="store_true", help="skip the compact records of the other BASELINE configs")
    ap.add_argument("--no-exchange", action="store_true", help="WORLD_SIZE > 1: leave the gradient exchange out of the step")
    ap.add_argument("--no-graph", action="store_true", help="launch the small workloads eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, spec, rank, world)
        return
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    run_ours(args, args.workload, spec, rank, world, local_rank)


if __name__ == "__main__":
    main()
