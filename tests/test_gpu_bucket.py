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
def test_peer_memory_allreduce_matches_nccl_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "bench_allreduce.py"), "--numel", "1376000", "--iters", "10"]  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["ok"], rec["notes"]
    assert rec["world"] == 2
