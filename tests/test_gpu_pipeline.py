"""Host staging helpers (``-m gpu``): the double-buffered pipeline must produce exactly the serial path's results."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import clip_path

pytestmark = pytest.mark.gpu


def test_prefetch_and_deferred_read_match_serial():
    import humanoid_amp_b200 as amp
    from humanoid_amp_b200.synthetic import skrl_style_discriminator_params

    dev = torch.device("cuda", 0)
    K, n, steps = 2, 20000, 6
    loader = amp.MotionLoader(clip_path("G1_walk"), dev)
    robot = amp.robot_for_clip(loader.dof_names)
    env = amp.AmpEnvPath(amp.AmpEnvCfg(motion_file=clip_path("G1_walk"), num_envs=1, num_amp_observations=K, robot=robot), dev,
                         motion_loader=loader)
    width = K * robot.amp_observation_space
    W, b = skrl_style_discriminator_params(width, seed=3, logit_gain=5.0)
    disc = amp.AmpDiscriminator(width, device=dev, max_rows=n)
    disc.load(W, b, torch.zeros(width, dtype=torch.float64), torch.ones(width, dtype=torch.float64))

    rng = np.random.default_rng(5)
    batches = [(rng.uniform(-0.1, float(loader.duration) + 0.1, n), np.zeros(n, dtype=np.int64)) for _ in range(steps)]

    serial = []
    for t, ids in batches:
        obs = env.collect_reference_motions(n, t, ids)
        serial.append(disc.style_reward(obs).view(-1).cpu())

    pre = amp.InputPrefetcher(dev, n, depth=2)
    reader = amp.ResultReader(dev, depth=2)
    obs_buf = torch.empty((n, width), device=dev)
    rew = [torch.empty(n, device=dev), torch.empty(n, device=dev)]
    got, prev = [], None
    slot = pre.submit(*batches[0])
    for i in range(steps):
        nxt = pre.submit(*batches[i + 1]) if i + 1 < steps else None
        t_d, i_d = pre.acquire(slot)
        o = env.collect_reference_motions(n, t_d, i_d, out=obs_buf)
        pre.release(slot)
        r = disc.style_reward(o, out=rew[i & 1])
        ticket = reader.read_async(r.view(-1))
        if prev is not None:
            got.append(prev.wait().clone())
        prev, slot = ticket, nxt
    got.append(prev.wait().clone())

    assert len(got) == steps
    for a, b_ in zip(serial, got):
        assert torch.equal(a, b_)  # same kernels, same inputs: bit-identical


def test_prefetcher_rejects_wrong_size():
    import humanoid_amp_b200 as amp

    pre = amp.InputPrefetcher("cuda:0", 16)
    with pytest.raises(ValueError):
        pre.submit(np.zeros(8), np.zeros(8, dtype=np.int64))
