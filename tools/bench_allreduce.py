"""Gradient all-reduce over peer memory (csrc/amp_bucket.cu) against NCCL, under torchrun on ONE node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_allreduce.py [--numel 2650000] [--iters 50]

Checks (every rank): the bucket result equals NCCL all_reduce(SUM) / world within fp32 summation-order noise, is bitwise
IDENTICAL on all ranks, stays correct over many back-to-back calls (epoch logic) and when the gradients are produced
in place by AmpDiscriminatorUpdate.  Times both with CUDA events (barrier + synchronize on both sides, max over ranks).
Rank 0 prints one JSON line ("ok": true/false).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=2_650_000)  # policy + value + discriminator of the G1 dance config
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--in-switch", choices=["auto", "0", "1"], default="auto",
                    help="AMP_B200_BUCKET_IN_SWITCH for the bucket under test: auto = from 4 ranks up, 1 = also at 2 ranks, 0 = never")
    a = ap.parse_args()
    if a.in_switch != "auto":
        os.environ["AMP_B200_BUCKET_IN_SWITCH"] = a.in_switch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import humanoid_amp_b200 as amp

    ok, notes = True, []
    bucket = amp.GradientBucket(a.numel, dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    # 1. correctness against NCCL, many calls back to back, odd sub-ranges
    for it, (off, cnt) in enumerate([(0, a.numel), (0, a.numel), (4, a.numel - 4), (1024, 1000), (0, 4), (8, (a.numel - 8) // 4 * 4), (0, a.numel)]):
        src = torch.randn(a.numel, device=dev, generator=g) * (1.0 + rank)
        bucket.flat[: a.numel].copy_(src)
        want = src.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        want /= world
        bucket.all_reduce_mean(off, cnt)
        got = bucket.flat[: a.numel]
        end = off + (cnt + 3) // 4 * 4  # the kernel works on whole quads
        err = float((got[off:end] - want[off:end]).abs().max())
        scale = float(want.abs().max())
        if not err <= 1e-5 * max(scale, 1.0):
            ok = False
            notes.append(f"call {it}: max err {err:.3e}")
        if not torch.equal(got[:off], src[:off]) or not torch.equal(got[end:], src[end:]):
            ok = False
            notes.append(f"call {it}: elements outside the range changed")
        mine = got[off:end].clone()
        ref0 = mine.clone()
        dist.broadcast(ref0, 0)
        if not torch.equal(mine, ref0):
            ok = False
            notes.append(f"call {it}: ranks disagree bitwise")
    if bucket.poll_status() != 0:
        ok = False
        notes.append("device status word set")
    try:  # a ragged range that ends INSIDE the bucket would average up to 3 floats it was not asked to touch: refused
        bucket.all_reduce_mean(0, 1001)
        ok = False
        notes.append("ragged interior range was accepted")
    except amp.AmpB200Error:
        pass

    # 1b. the all-reduce inside a CUDA graph: the barrier epochs are device state, so every replay uses fresh ones
    src = torch.randn(a.numel, device=dev, generator=g) * (1.0 + rank)
    staged = src.clone()

    def graph_body():
        bucket.flat[: a.numel].copy_(staged)
        bucket.all_reduce_mean(0, a.numel)

    graph = amp.capture_step(graph_body, dev, warmup=2)
    want = src.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    want /= world
    for rep in range(4):
        graph.replay()
        torch.cuda.synchronize(dev)
        err = float((bucket.flat[: a.numel] - want).abs().max())
        if not err <= 1e-5 * max(1.0, float(want.abs().max())):
            ok = False
            notes.append(f"graph replay {rep}: max err {err:.3e}")
    bucket.all_reduce_mean(0, a.numel)  # and eager calls still interleave with replays
    if bucket.poll_status() != 0:
        ok = False
        notes.append("device status word set after graph replays")

    # 2. gradients produced in place by the discriminator update
    in_features, hidden, B = 166, (1024, 512), 512
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    W, b = skrl_style_discriminator_params(in_features, seed=42)
    W, b = [w.to(dev) for w in W], [x.to(dev) for x in b]
    shapes = [tuple(w.shape) for w in W] + [tuple(x.shape) for x in b]
    disc_bucket = amp.GradientBucket(sum(w.numel() for w in W) + sum(x.numel() for x in b), dev)
    views = disc_bucket.carve(shapes)
    upd = amp.AmpDiscriminatorUpdate(in_features, hidden, max_batch_rows=B, device=dev)
    batches = [torch.randn(B, in_features, device=dev, generator=g).clamp_(-5, 5) for _ in range(3)]
    upd(W, b, *batches, grad_weights=views[:3], grad_biases=views[3:])
    local_grads = disc_bucket.flat[: disc_bucket.numel].clone()
    want = local_grads.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    want /= world
    disc_bucket.all_reduce_mean()
    err = float((disc_bucket.flat[: disc_bucket.numel] - want).abs().max())
    if not err <= 1e-5 * max(1.0, float(want.abs().max())):
        ok = False
        notes.append(f"in-place discriminator gradients: max err {err:.3e}")

    # 3. timing
    scratch = torch.randn(a.numel, device=dev, generator=g)

    def nccl():
        dist.all_reduce(scratch, op=dist.ReduceOp.SUM)
        scratch.div_(world)

    ms_nccl = timed(nccl, a.iters, dev)
    ms_bucket = timed(lambda: bucket.all_reduce_mean(0, a.numel), a.iters, dev)
    phases = bucket.last_timing_us()
    if bucket.poll_status() != 0:
        ok = False
        notes.append("device status word set after timing")
    # the other form of the bucket (peer memory when the default came up in the switch): same checks in short, and its time
    ms_other = None
    if bucket.in_switch:
        before = os.environ.get("AMP_B200_BUCKET_IN_SWITCH")
        os.environ["AMP_B200_BUCKET_IN_SWITCH"] = "0"
        other = amp.GradientBucket(a.numel, dev)
        if before is None:
            del os.environ["AMP_B200_BUCKET_IN_SWITCH"]
        else:
            os.environ["AMP_B200_BUCKET_IN_SWITCH"] = before
        src = torch.randn(a.numel, device=dev, generator=g) * (1.0 + rank)
        other.flat[: a.numel].copy_(src)
        want = src.clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        want /= world
        other.all_reduce_mean(0, a.numel)
        err = float((other.flat[: a.numel] - want).abs().max())
        if other.in_switch or not err <= 1e-5 * max(1.0, float(want.abs().max())):
            ok = False
            notes.append(f"peer-memory form: in_switch={other.in_switch}, max err {err:.3e}")
        ms_other = timed(lambda: other.all_reduce_mean(0, a.numel), a.iters, dev)
        other.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        bytes_moved = 2.0 * (world - 1) / world * a.numel * 4
        print(json.dumps({"what": "gradient all-reduce (mean) of one flat fp32 bucket", "world": world, "numel": a.numel,
                          "ok": bool(flag.item()), "notes": notes, "in_switch": bucket.in_switch, "ms_nccl_allreduce_plus_div": ms_nccl,
                          "ms_bucket_kernel": ms_bucket, "ms_peer_memory_form": ms_other if bucket.in_switch else ms_bucket,
                          "speedup": ms_nccl / ms_bucket, "phases_us_rank0_last_call": phases, "nvlink_gbs_per_rank_each_way": bytes_moved / ms_bucket / 1e6}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
