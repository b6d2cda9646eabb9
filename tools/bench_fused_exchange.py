"""Discriminator update with the gradient exchange fused into its last kernel (amp_disc_train_step_exchange) against the same
update followed by the bucket all-reduce, under torchrun on ONE node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/bench_fused_exchange.py [--in-features 166 830] [--batch 4096] [--iters 20]

Checks (every rank, every width): the fused step leaves what step + all_reduce_mean leaves and what NCCL all_reduce(SUM) / world
makes of the local gradients (to the run-to-run noise of the step itself: its head / column-sum kernels accumulate with fp32
atomics, so two runs of the SAME step differ in the last bits), the ranks agree bit for bit after every call, the floats of the
bucket around the gradient range are untouched, repeated calls and CUDA-graph replays stay correct and interleave with plain
all-reduces on the same bucket.  Times both forms as CUDA graphs (events, barrier + synchronize on both sides, max over
ranks).  Rank 0 prints one JSON line ("ok": true/false).
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

from tools.bench_allreduce import timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-features", type=int, nargs="+", default=[166, 830])
    ap.add_argument("--batch", type=int, default=4096)  # agents/skrl_g1_dance_amp_cfg.yaml:94 discriminator_batch_size
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--in-switch", choices=["auto", "0", "1"], default="auto", help="AMP_B200_BUCKET_IN_SWITCH for the buckets (see bench_allreduce.py)")
    a = ap.parse_args()
    if a.in_switch != "auto":
        os.environ["AMP_B200_BUCKET_IN_SWITCH"] = a.in_switch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    ok, notes, rows = True, [], []

    def bad(msg):
        nonlocal ok
        ok = False
        notes.append(msg)

    for inf in a.in_features:
        W, b = skrl_style_discriminator_params(inf, seed=42)  # the replicas hold the same parameters
        W, b = [w.to(dev) for w in W], [x.to(dev) for x in b]
        n_disc = sum(w.numel() for w in W) + sum(x.numel() for x in b)
        # torch parameter order, 12 floats of "someone else's gradients" in front and 7 behind
        bucket = amp.GradientBucket(12 + n_disc + 7, dev)
        views = bucket.carve([(12,), W[0].shape, b[0].shape, W[1].shape, b[1].shape, W[2].shape, b[2].shape, (7,)])
        gW, gb = [views[1], views[3], views[5]], [views[2], views[4], views[6]]
        upd = amp.AmpDiscriminatorUpdate(inf, (1024, 512), max_batch_rows=a.batch, device=dev)
        g = torch.Generator(device=dev).manual_seed(99 + rank)  # every rank has its own batches
        batches = [torch.randn(a.batch, inf, device=dev, generator=g).clamp_(-5, 5) * (1.0 + 0.1 * rank) for _ in range(3)]

        def plain():
            upd(W, b, *batches, grad_weights=gW, grad_biases=gb)
            bucket.all_reduce_mean(12, n_disc + 7)  # the range ends at the bucket's end (ragged counts are only accepted there)

        def fused():
            upd(W, b, *batches, grad_weights=gW, grad_biases=gb, bucket=bucket)

        # reference: local gradients -> NCCL
        upd(W, b, *batches, grad_weights=gW, grad_biases=gb)
        want = bucket.flat[12 : 12 + n_disc].clone()
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        want /= world
        bucket.flat.fill_(3.0)
        plain()
        two_launch = bucket.flat[12 : 12 + n_disc].clone()
        tol = 1e-5 * max(1.0, float(want.abs().max()))

        def agree(what):
            got = bucket.flat[12 : 12 + n_disc]
            err = float((got - two_launch).abs().max())
            if not err <= tol:
                bad(f"in={inf} {what}: fused step differs from step + all-reduce by {err:.3e}")
            ref0 = got.clone()
            dist.broadcast(ref0, 0)
            if not torch.equal(ref0, got):
                bad(f"in={inf} {what}: ranks disagree bitwise")

        for rep in range(3):
            bucket.flat.fill_(5.0)
            fused()
            agree(f"call {rep}")
            if float(bucket.flat[:12].min()) != 5.0 or float(bucket.flat[:12].max()) != 5.0 or not bool((bucket.flat[12 + n_disc :][:7] == 5.0).all()):
                bad(f"in={inf} call {rep}: floats outside the gradient range changed")
            if rep == 1:
                plain()  # the two forms share the bucket's flags and epochs
        err = float((two_launch - want).abs().max())
        if not err <= tol:
            bad(f"in={inf}: mean differs from NCCL by {err:.3e}")
        ref0 = two_launch.clone()
        dist.broadcast(ref0, 0)
        if not torch.equal(ref0, two_launch):
            bad(f"in={inf}: ranks disagree bitwise")

        g_plain = amp.capture_step(plain, dev, warmup=2)
        g_fused = amp.capture_step(fused, dev, warmup=2)
        for rep in range(3):
            bucket.flat.fill_(1.0)
            g_fused.replay()
            torch.cuda.synchronize(dev)
            agree(f"graph replay {rep}")
            g_plain.replay()
        if bucket.poll_status() != 0:
            bad(f"in={inf}: device status word set")
        ms_plain = timed(g_plain.replay, a.iters, dev)
        ms_fused = timed(g_fused.replay, a.iters, dev)
        for _ in range(5):  # a few replays put the ranks in lock step before the stamps are read
            g_fused.replay()
        phases = bucket.last_timing_us()
        rows.append({"in_features": inf, "batch_rows": a.batch, "gradient_floats": n_disc, "all_reduce_in_switch": bucket.in_switch,
                     "us_step_plus_allreduce": ms_plain * 1e3,
                     "us_fused_step": ms_fused * 1e3, "phases_us_rank0_last_fused_call": phases})
        del g_plain, g_fused
        upd.close()
        bucket.close()

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "discriminator update with the gradient exchange fused into its last kernel vs update + all-reduce (CUDA graphs)",
                          "world": world, "ok": bool(flag.item()), "ranks_bit_identical": bool(flag.item()), "notes": notes, "widths": rows}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
