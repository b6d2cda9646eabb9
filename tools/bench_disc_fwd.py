#!/usr/bin/env python
"""Developer micro-benchmark of the fused discriminator forward (amp_disc_style_reward) alone.

    python tools/bench_disc_fwd.py [--rows 1000000] [--in-features 166] [--iters 20] [--pair] [--sustained-s 0]

Prints one JSON line per configuration: ms per call (CUDA events on the launching stream, after warm-up) and the algorithmic
TFLOP/s.  With an AMP_DISC_PROFILE=1 build of the library every call also dumps the in-kernel wait counters to stderr.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[1_000_000])
    ap.add_argument("--in-features", type=int, nargs="+", default=[166])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pair", action="store_true", help="(default) CTA-pair kernel")
    ap.add_argument("--single", action="store_true", help="single-CTA kernel")
    ap.add_argument("--graph", action="store_true", help="replay each call as a CUDA graph (small batches)")
    ap.add_argument("--warm", action="store_true", help="do NOT flush L2 between calls of a small batch (weights / rows stay resident)")
    args = ap.parse_args()
    os.environ["AMP_B200_DISC_PAIR"] = "0" if args.single else "1"
    args.pair = not args.single
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for inf in args.in_features:
        W, b = skrl_style_discriminator_params(inf, seed=42, logit_gain=5.0)
        for M in args.rows:
            disc = amp.AmpDiscriminator(inf, device=dev, max_rows=M)
            disc.load(W, b, torch.zeros(inf, dtype=torch.float64), torch.ones(inf, dtype=torch.float64))
            x = torch.randn(M, inf, device=dev)
            out = torch.empty(M, device=dev)
            call = lambda: disc.style_reward(x, out=out)  # noqa: E731
            for _ in range(args.warmup):
                call()
            torch.cuda.synchronize()
            small = M * inf * 4 < (64 << 20) and not args.warm
            g = amp.capture_step(call, dev) if args.graph else None
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
            for a, c in evs:
                if small:
                    flush.zero_()
                a.record()
                g.replay() if g is not None else call()
                c.record()
            torch.cuda.synchronize()
            ms = sorted(a.elapsed_time(c) for a, c in evs)
            med = ms[len(ms) // 2]
            flops = M * 2.0 * (inf * 1024 + 1024 * 512 + 512)
            print(json.dumps({"in_features": inf, "rows": M, "pair": args.pair, "graph": args.graph, "ms_median": med, "ms_min": ms[0],
                              "tflops": flops / (med * 1e-3) / 1e12, "l2_flushed_between_calls": small}), flush=True)
            del disc


if __name__ == "__main__":
    main()
