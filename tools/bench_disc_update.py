"""Time the discriminator loss + gradient step on one B200 and put it next to what a skrl user runs today:
the same loss through torch autograd on the same GPU (fp32, TF32 off = torch's default; and bf16 autocast).

    python tools/bench_disc_update.py [--in-features 830] [--batch 4096] [--steps 50] [--no-torch]

Algorithmic flops of one step (B rows per source; derivation in oracle/disc_train_oracle.py):
    forward + backward of 3B rows: 6B (2 in h1 + 3 h1 h2);  gradient penalty on B rows: 6B (in h1 + h1 h2)
    total = 6 B (3 in h1 + 4 h1 h2)                     = 114.2 GFLOP at B = 4096, in = 830, 1024-512
Prints one JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_cuda(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / steps


def torch_step(W, b, agent, replay, motion, autocast):
    """The literal skrl expression on the GPU (what runs today); returns the loss after backward."""
    for p in W + b:
        p.grad = None
    motion = motion.detach().requires_grad_(True)

    def mlp(x):
        h = torch.relu(torch.nn.functional.linear(x, W[0], b[0]))
        h = torch.relu(torch.nn.functional.linear(h, W[1], b[1]))
        return torch.nn.functional.linear(h, W[2], b[2])

    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits, replay_logits, motion_logits = mlp(agent), mlp(replay), mlp(motion)
        cat = torch.cat([logits, replay_logits], 0).float()
        motion_logits = motion_logits.float()
        bce = torch.nn.BCEWithLogitsLoss()
        loss = 0.5 * (bce(cat, torch.zeros_like(cat)) + bce(motion_logits, torch.ones_like(motion_logits)))
        loss = loss + 0.05 * torch.sum(torch.square(torch.flatten(W[2])))
        grad = torch.autograd.grad(motion_logits, motion, grad_outputs=torch.ones_like(motion_logits), create_graph=True,
                                   retain_graph=True, only_inputs=True)[0]
        loss = loss + 5.0 * torch.sum(torch.square(grad.float()), dim=-1).mean()
        loss = loss + 1.0e-4 * torch.sum(torch.square(torch.cat([torch.flatten(w) for w in W], dim=-1)))
        loss = loss * 5.0
    loss.backward()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-features", type=int, default=830)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-torch", action="store_true")
    a = ap.parse_args()

    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    dev = "cuda:0"
    h1, h2, B, n_in = 1024, 512, a.batch, a.in_features
    W, b = skrl_style_discriminator_params(n_in, seed=42)
    W = [w.to(dev) for w in W]
    b = [x.to(dev) for x in b]
    g = torch.Generator(device=dev).manual_seed(1)
    raw = [torch.randn(B, n_in, device=dev, generator=g) * 1.5 + 0.2 for _ in range(3)]
    scaler = amp.RunningStandardScaler(n_in, device=dev)
    upd = amp.AmpDiscriminatorUpdate(n_in, (h1, h2), max_batch_rows=B, device=dev)
    gW = [torch.empty_like(w) for w in W]
    gb = [torch.empty_like(x) for x in b]

    flops = 6.0 * B * (3.0 * n_in * h1 + 4.0 * h1 * h2)
    ms_full = time_cuda(lambda: upd(W, b, *raw, scaler=scaler, train=True, grad_weights=gW, grad_biases=gb), a.steps, a.warmup)

    def staged_only():
        for i in range(3):
            upd.stage(i, raw[i], scaler, train=False)
        upd.loss_and_grads(W, b, gW, gb)

    ms_eval_scaler = time_cuda(staged_only, a.steps, a.warmup)
    graph = amp.capture_step(staged_only, dev)
    ms_graph = time_cuda(graph.replay, a.steps, a.warmup)
    out = {
        "what": "discriminator loss + gradients (skrl AMP._update block), one step",
        "config": {"in_features": n_in, "hidden": [h1, h2], "rows_per_source": B, "sources": 3},
        "algorithmic_gflop": flops / 1e9,
        "ms_step_with_scaler_update": ms_full,
        "ms_step_scaler_eval": ms_eval_scaler,
        "ms_step_scaler_eval_cuda_graph": ms_graph,
        "tflops_step_with_scaler_update": flops / ms_full / 1e9,
        "tflops_step_scaler_eval": flops / ms_eval_scaler / 1e9,
    }
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["bf16_tflops"]
        out["tensor_peak_burst_tflops"] = peak
        out["frac_of_burst_peak"] = out["tflops_step_scaler_eval"] / peak
    except Exception:
        pass
    if not a.no_torch:
        Wt = [w.clone().requires_grad_(True) for w in W]
        bt = [x.clone().requires_grad_(True) for x in b]
        normed = [scaler(x) for x in raw]
        for name, autocast in (("torch_autograd_fp32", False), ("torch_autograd_bf16_autocast", True)):
            ms = time_cuda(lambda: torch_step(Wt, bt, *normed, autocast), max(5, a.steps // 5), 3)
            out[f"ms_{name}"] = ms
            out[f"speedup_vs_{name}"] = ms / ms_eval_scaler
    print(json.dumps(out))


if __name__ == "__main__":
    main()
