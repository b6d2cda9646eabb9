import os, sys, torch
sys.path.insert(0, '/root/repo')
import humanoid_amp_b200 as amp
from humanoid_amp_b200.synthetic import skrl_style_discriminator_params
for pair in ("1", "0"):
    os.environ["AMP_B200_DISC_PAIR"] = pair
    for inf, M in ((166, 148 * 128 * 3 + 77), (166, 1000), (83, 40000), (830, 3000)):
        W, b = skrl_style_discriminator_params(inf, seed=1, logit_gain=3.0)
        d = amp.AmpDiscriminator(inf, device="cuda:0", max_rows=2048)
        d.load(W, b, torch.zeros(inf, dtype=torch.float64), torch.ones(inf, dtype=torch.float64))
        x = torch.randn(M, inf, device="cuda")
        r = d.style_reward(x)
        idx = torch.randint(0, M, (5000,), device="cuda")
        r2 = d.style_reward_sampled(x, idx)
        torch.cuda.synchronize()
        assert torch.equal(r2, r[idx]) or (r2 - r[idx]).abs().max() < 1e-6, (pair, inf, M)
        print("ok", pair, inf, M, float(r.mean()))
