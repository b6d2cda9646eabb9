"""world_size-2 ``gloo`` tests (CPU) of the multi-GPU plumbing: flat gradient all-reduce and env sharding.

The data path itself (sampling, AMP obs, style reward) needs no collective: every rank owns its env shard and a full
clip replica.  The only exchange is skrl's ``Model.reduce_parameters`` (SUM / world) restated in
``humanoid_amp_b200/distributed.py``.
"""

from __future__ import annotations

import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from humanoid_amp_b200.distributed import reduce_parameters, shard_envs

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical initial parameters on every rank (skrl broadcasts them from rank 0)
        policy = torch.nn.Linear(7, 5)
        value = torch.nn.Linear(7, 1)
        disc = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.ReLU(), torch.nn.Linear(4, 1))
        frozen = torch.nn.Parameter(torch.ones(3))  # never receives a gradient: must be treated as zeros and left None
        g = torch.Generator().manual_seed(100 + rank)  # different data per rank = different env shard
        (policy(torch.randn(9, 7, generator=g)).sum() + value(torch.randn(9, 7, generator=g)).pow(2).sum()).backward()
        disc(torch.randn(9, 6, generator=g)).sum().backward()
        groups = [list(policy.parameters()), list(value.parameters()), list(disc.parameters()) + [frozen]]
        local = [p.grad.clone() if p.grad is not None else None for grp in groups for p in grp]
        flat = reduce_parameters(groups)
        flat2 = reduce_parameters(groups, flat=flat)  # second call reuses the buffer and averages again (idempotent on equal grads)
        assert flat2.data_ptr() == flat.data_ptr()
        after = [p.grad.clone() if p.grad is not None else None for grp in groups for p in grp]
        begin, end = shard_envs(4097, rank, world)
        torch.save({"local": local, "after": after, "shard": (begin, end), "numel": flat.numel()}, os.path.join(result_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_reduce_parameters_world2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    n_params = len(r[0]["local"])
    assert r[0]["numel"] == sum(g.numel() for g in r[0]["local"] if g is not None) + 3
    for i in range(n_params):
        if r[0]["local"][i] is None:
            assert r[0]["after"][i] is None and r[1]["after"][i] is None
            continue
        mean = (r[0]["local"][i] + r[1]["local"][i]) / world
        for rank in range(world):
            assert torch.allclose(r[rank]["after"][i], mean, rtol=0, atol=1e-6), f"param {i} rank {rank}"
    differing = [i for i in range(n_params) if r[0]["local"][i] is not None and not torch.equal(r[0]["local"][i], r[1]["local"][i])]
    assert len(differing) >= 4  # the ranks really had different gradients (bias grads of a plain sum coincide)
    assert r[0]["shard"] == (0, 2049) and r[1]["shard"] == (2049, 4097)


def _fd_worker(rank, world, port, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from humanoid_amp_b200.distributed import swap_fds

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank owns a pipe whose write end travels to the others; rank 0 a second one ("the multicast object")
        r_own, w_own = os.pipe()
        r_root, w_root = os.pipe() if rank == 0 else (-1, -1)
        fds, root = swap_fds(w_own, w_root, rank, world)
        assert fds[rank] == -1 and all(fd >= 0 for p, fd in enumerate(fds) if p != rank)
        assert (root == -1) == (rank == 0)
        for p, fd in enumerate(fds):  # write into every peer's pipe through the descriptor that was handed over
            if p != rank:
                os.write(fd, bytes([rank]))
                os.close(fd)
        if rank != 0:
            os.write(root, bytes([100 + rank]))
            os.close(root)
        dist.barrier()
        os.close(w_own)
        own = sorted(os.read(r_own, 64))
        extra = sorted(os.read(r_root, 64)) if rank == 0 else None
        if rank == 0:
            os.close(w_root)
        with open(os.path.join(result_dir, f"fd{rank}.json"), "w") as f:
            json.dump({"own": own, "extra": extra}, f)
    finally:
        dist.destroy_process_group()


def test_swap_fds_world3_gloo(tmp_path):
    """The descriptor exchange behind the shared gradient bucket (SCM_RIGHTS over abstract unix sockets): every rank must end up
    with a working descriptor for every other rank's object and for rank 0's extra one."""
    world, port = 3, _free_port()
    mp.spawn(_fd_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        rec = json.load(open(tmp_path / f"fd{rank}.json"))
        assert rec["own"] == [p for p in range(world) if p != rank]
        assert rec["extra"] == ([101, 102] if rank == 0 else None)


def test_reduce_parameters_is_noop_without_process_group():
    from humanoid_amp_b200.distributed import reduce_parameters

    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    before = lin.weight.grad.clone()
    assert reduce_parameters([list(lin.parameters())]) is None
    assert torch.equal(lin.weight.grad, before)


def test_bench_reference_arm_under_torchrun_only_rank0_prints(tmp_path):
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints one JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--impl", "reference", "--workload", "g1_walk_4096x2"]  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["cpu_baseline"]["kind"] in ("reference", "port")
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["value"] > 0 and rec["unit"] == "samples/s"
