"""The gradient bucket (``-m gpu``): single-rank behaviour on one GPU, and -- when the box has
two GPUs -- the peer-memory all-reduce against NCCL under a 2-rank ``torchrun`` (``tools/bench_allreduce.py``)."""

from __future__ import annotations

import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_single_rank_bucket_is_plain_device_memory():
    import humanoid_amp_b200 as amp

    bucket = amp.GradientBucket(1001, DEV)
    assert bucket.world == 1 and bucket.capacity == 1004 and bucket.flat.device.type == "cuda"
    assert float(bucket.flat.abs().sum()) == 0.0  # zero initialised
    a, b = bucket.carve([(10, 50), (501,)])
    a.fill_(1.5)
    b.fill_(-2.0)
    assert float(bucket.flat[:1001].sum()) == 1.5 * 500 - 2.0 * 501
    before = bucket.flat.clone()
    bucket.all_reduce_mean()  # world of one: identity
    assert torch.equal(bucket.flat, before) and bucket.poll_status() == 0
    with pytest.raises(ValueError):
        bucket.carve([(1000,), (2,)])
    with pytest.raises(amp.AmpB200Error):
        bucket.all_reduce_mean(2, 8)  # offset not 16-byte aligned
    with pytest.raises(amp.AmpB200Error):
        bucket.all_reduce_mean(0, 2000)


def test_discriminator_gradients_land_in_the_bucket():
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    W, b = skrl_style_discriminator_params(166, seed=1)
    shapes = [tuple(w.shape) for w in W] + [tuple(x.shape) for x in b]
    bucket = amp.GradientBucket(sum(w.numel() for w in W) + sum(x.numel() for x in b), DEV)
    views = bucket.carve(shapes)
    upd = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=256, device=DEV)
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn(256, 166, generator=g).clamp_(-5, 5).to(DEV) for _ in range(3)]
    _, gW, gb = upd(W, b, *batches, grad_weights=views[:3], grad_biases=views[3:])
    _, gW2, gb2 = upd(W, b, *batches)
    for v, ref in zip(views, gW2 + gb2):
        assert v.data_ptr() >= bucket.flat.data_ptr() and torch.allclose(v, ref, rtol=1e-4, atol=1e-7)
    assert float(bucket.flat.abs().sum()) > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box (run with gpurun --gpus 2)")
@pytest.mark.parametrize("in_switch", ["auto", "1"])
def test_allreduce_matches_nccl_world2(in_switch):
    """Both forms of the bucket under a 2-rank torchrun: peer memory (the default below 4 ranks) and, forced, the shared form
    whose all-reduce runs in the NVSwitch (skipped where the node has no multicast support: the tool then reports the fallback)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "bench_allreduce.py"), "--numel", "1376000", "--iters", "10",
           "--in-switch", in_switch]  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["ok"], rec["notes"]
    assert rec["world"] == 2
    if in_switch == "auto":
        assert not rec["in_switch"]
    elif not rec["in_switch"]:
        pytest.skip("no NVSwitch multicast on this node: the bucket fell back to peer memory")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box (run with gpurun --gpus 2)")
def test_second_device_from_one_process():
    """One process driving two GPUs (the current device stays cuda:0): every entry point must launch on the tensor's device
    -- the shims call amp_set_device before each call, kernel attributes are configured per device."""
    import numpy as np

    import humanoid_amp_b200 as amp
    from conftest import clip_path
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    results = []
    for dev in ("cuda:0", "cuda:1"):
        clip = clip_path("G1_walk")
        env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=clip, num_envs=64, num_amp_observations=2, robot=amp.G1), dev)
        rng = np.random.default_rng(0)
        times = rng.uniform(0, 0.6, 2048)
        obs = env.collect_reference_motions(2048, times, np.zeros(2048, dtype=np.int64))
        W, b = skrl_style_discriminator_params(166, seed=3)
        disc = amp.AmpDiscriminator(166, device=dev, max_rows=2048)
        disc.load(W, b, torch.zeros(166, dtype=torch.float64), torch.ones(166, dtype=torch.float64))
        reward = disc.style_reward(obs)
        upd = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=512, device=dev)
        _, gW, gb = upd(W, b, obs[:512], obs[512:1024], obs[1024:1536])
        assert obs.device == torch.device(dev) and reward.device == torch.device(dev) and gW[0].device == torch.device(dev)
        results.append((obs.cpu(), reward.cpu(), gW[0].cpu(), gb[0].cpu()))
    assert torch.cuda.current_device() == 0
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])
    assert torch.allclose(results[0][2], results[1][2], rtol=1e-4, atol=1e-7)
    assert torch.allclose(results[0][3], results[1][3], rtol=1e-4, atol=1e-7)


def test_fused_exchange_step_placement_rules_and_world_of_one():
    """``loss_and_grads(bucket=...)``: on one rank it is the plain step; the six gradient views must tile one quad-aligned range
    of the bucket with dL/db3 last -- anything else is refused (the same rules on every world size)."""
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    W, b = skrl_style_discriminator_params(166, seed=1)
    W, b = [w.to(DEV) for w in W], [x.to(DEV) for x in b]
    n = sum(w.numel() for w in W) + sum(x.numel() for x in b)
    upd = amp.AmpDiscriminatorUpdate(166, (1024, 512), max_batch_rows=256, device=DEV)
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn(256, 166, generator=g).clamp_(-5, 5).to(DEV) for _ in range(3)]
    _, gW_ref, gb_ref = upd(W, b, *batches)

    # torch parameter order (W1, b1, W2, b2, W3, b3), behind 8 floats of someone else's gradients
    bucket = amp.GradientBucket(n + 8 + 5, DEV)
    bucket.flat.fill_(7.0)
    views = bucket.carve([(8,), W[0].shape, b[0].shape, W[1].shape, b[1].shape, W[2].shape, b[2].shape, (5,)])
    upd(W, b, *batches, grad_weights=[views[1], views[3], views[5]], grad_biases=[views[2], views[4], views[6]], bucket=bucket)
    # (two runs of the step differ in the last bits: the head / column-sum kernels accumulate with fp32 atomics)
    for v, ref in zip([views[1], views[3], views[5], views[2], views[4], views[6]], gW_ref + gb_ref):
        assert torch.allclose(v, ref, rtol=1e-4, atol=1e-6)
    assert float(views[0].min()) == 7.0 and float(views[7].min()) == 7.0  # neighbours untouched
    # weights-then-biases order is fine too
    v2 = bucket.carve([s.shape for s in W] + [x.shape for x in b])
    upd(W, b, *batches, grad_weights=v2[:3], grad_biases=v2[3:], bucket=bucket)
    for v, ref in zip(v2, gW_ref + gb_ref):
        assert torch.allclose(v, ref, rtol=1e-4, atol=1e-6)
    # refused: not in the bucket / a gap / db3 not last / range not on a quad
    with pytest.raises(amp.AmpB200Error, match="not inside the bucket"):
        upd(W, b, *batches, grad_weights=[torch.empty_like(w) for w in W], grad_biases=v2[3:], bucket=bucket)
    gap = bucket.carve([W[0].shape, (4,), W[1].shape, W[2].shape, b[0].shape, b[1].shape, b[2].shape])
    with pytest.raises(amp.AmpB200Error, match="side by side"):
        upd(W, b, *batches, grad_weights=[gap[0], gap[2], gap[3]], grad_biases=gap[4:], bucket=bucket)
    b3_first = bucket.carve([(4,), b[2].shape, (3,), W[0].shape, W[1].shape, W[2].shape, b[0].shape, b[1].shape])
    with pytest.raises(amp.AmpB200Error):
        upd(W, b, *batches, grad_weights=b3_first[3:6], grad_biases=[b3_first[6], b3_first[7], b3_first[1]], bucket=bucket)
    odd = bucket.carve([(2,)] + [s.shape for s in W] + [x.shape for x in b])
    with pytest.raises(amp.AmpB200Error, match="16-byte"):
        upd(W, b, *batches, grad_weights=odd[1:4], grad_biases=odd[4:], bucket=bucket)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on the box (run with gpurun --gpus 2)")
def test_fused_exchange_step_matches_step_plus_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "bench_fused_exchange.py"), "--iters", "10"]  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["ok"], rec["notes"]
    assert rec["world"] == 2 and rec["ranks_bit_identical"]
