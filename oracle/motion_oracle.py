"""Oracle restatement of the reference motion library (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

Follows ``/root/reference/motions/motion_loader.py``:

* clip loading / concatenation ............ ``:98-164``
* frame index + blend (float64, host) ..... ``:281-307``
* linear interpolation .................... ``:186-215``
* shortest-arc slerp with two fallbacks ... ``:217-279``
* ``sample_times`` / ``sample`` ............ ``:309-390``

The arithmetic is kept in the reference's operation order (every torch op rounds separately to fp32; numpy index math is
float64 with round-half-even) because two hard thresholds in the slerp make the result sensitive to 1-ulp changes.
Pinned bit-for-bit against the live reference in ``tests/test_oracle_pins.py``.
"""

from __future__ import annotations

import numpy as np
import torch

_FIELDS = (
    "dof_positions",
    "dof_velocities",
    "body_positions",
    "body_rotations",
    "body_linear_velocities",
    "body_angular_velocities",
)


def frame_blend_f64(times, motion_ids, durations, traj_starts, traj_ends, dt):
    """Reference ``_compute_frame_blend`` (``motion_loader.py:294-305``).

    phase  = clip(t / duration, 0, 1)
    i0     = rint(phase * (end - start))          (round-half-even, NOT floor -> blend in [-0.5, 0.5] inside the clip)
    i1     = min(i0 + 1, end - start)
    blend  = round((t - i0 * dt) / dt, 5)          (numpy: multiply by 1e5, rint, divide by 1e5)
    Returns global frame indices (start + local) as int64 and the float64 blend.
    """
    times = np.asarray(times)
    motion_ids = np.asarray(motion_ids)
    dur = durations[motion_ids]
    first = traj_starts[motion_ids]
    span = traj_ends[motion_ids] - first
    phase = np.clip(times / dur, 0.0, 1.0)
    local0 = (phase * span).round(decimals=0).astype(int)
    local1 = np.minimum(local0 + 1, span)
    blend = ((times - local0 * dt) / dt).round(decimals=5)
    return first + local0, first + local1, blend


def _broadcast_blend(blend: torch.Tensor, ndim: int) -> torch.Tensor:
    # reference :211-214 / :242-245 -- one trailing unsqueeze per data dimension beyond the first
    for _ in range(max(ndim - 1, 0)):
        blend = blend.unsqueeze(-1)
    return blend


def lerp_f32(a: torch.Tensor, b: torch.Tensor, blend: torch.Tensor) -> torch.Tensor:
    """Reference ``_interpolate`` body (``motion_loader.py:215``): ``(1.0 - blend) * a + blend * b``."""
    w = _broadcast_blend(blend, a.ndim)
    return (1.0 - w) * a + w * b


def slerp_f32(q0: torch.Tensor, q1: torch.Tensor, blend: torch.Tensor) -> torch.Tensor:
    """Reference ``_slerp`` body (``motion_loader.py:247-279``), wxyz, output not renormalised.

    Order of evaluation that must be preserved:
      c  = ((w0*w1 + x0*x1) + y0*y1) + z0*z1            (four separately rounded products)
      q1 = -q1 where c < 0 ;  c = |c|
      h  = acos(c) ; s = sqrt(1.0 - c*c)
      ra = sin((1 - blend) * h) / s ; rb = sin(blend * h) / s
      out = ra*q0 + rb*q1
      out = 0.5*q0 + 0.5*q1   where |s| < 0.001          (ignores blend)
      out = q0                where |c| >= 1             (also masks the NaNs produced by acos(c > 1))
    """
    w = _broadcast_blend(blend, q0.ndim)
    c = q0[..., 0] * q1[..., 0] + q0[..., 1] * q1[..., 1] + q0[..., 2] * q1[..., 2] + q0[..., 3] * q1[..., 3]
    flip = c < 0
    q1 = q1.clone()
    q1[flip] = -q1[flip]
    c = torch.abs(c).unsqueeze(-1)

    half = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    ra = torch.sin((1 - w) * half) / s
    rb = torch.sin(w * half) / s

    # components are assembled x, y, z, w in the reference and concatenated as w, x, y, z; the per-component value is the
    # same expression either way
    parts = [ra * q0[..., i : i + 1] + rb * q1[..., i : i + 1] for i in range(4)]
    out = torch.cat(parts, dim=-1)
    out = torch.where(torch.abs(s) < 0.001, 0.5 * q0 + 0.5 * q1, out)
    out = torch.where(torch.abs(c) >= 1, q0, out)
    return out


class OracleMotionLoader:
    """CPU restatement of the reference ``MotionLoader`` working on already-resolved ``.npz`` paths."""

    def __init__(self, files, device="cpu"):
        if isinstance(files, (str, bytes)):
            files = [files]
        files = list(files)
        if not files:
            raise ValueError("no motion files")
        self.device = device
        chunks = {k: [] for k in _FIELDS}
        starts, ends, durs = [], [], []
        cursor = 0
        self.dt = None
        for path in files:
            with np.load(path) as data:
                if self.dt is None:  # names and dt come from the FIRST file only (reference :119-122)
                    self._dof_names = data["dof_names"].tolist()
                    self._body_names = data["body_names"].tolist()
                    self.dt = 1.0 / data["fps"]
                for k in _FIELDS:
                    chunks[k].append(data[k])
                n = data["dof_positions"].shape[0]
            starts.append(cursor)
            cursor += n
            ends.append(cursor - 1)
            durs.append(self.dt * (n - 1))
        self.traj_starts = np.array(starts)
        self.traj_ends = np.array(ends)
        self.durations = np.array(durs)
        self.num_trajectories = len(files)
        self.num_frames = cursor
        self.duration = float(np.sum(self.durations))
        for k in _FIELDS:
            setattr(self, k, torch.tensor(np.concatenate(chunks[k]), dtype=torch.float32, device=device))

    # -- metadata ----------------------------------------------------------------------------------------------------
    @property
    def dof_names(self):
        return self._dof_names

    @property
    def body_names(self):
        return self._body_names

    @property
    def num_dofs(self):
        return len(self._dof_names)

    @property
    def num_bodies(self):
        return len(self._body_names)

    def get_dof_index(self, names):
        out = []
        for n in names:
            assert n in self._dof_names, f"The specified DOF name ({n}) doesn't exist: {self._dof_names}"
            out.append(self._dof_names.index(n))
        return out

    def get_body_index(self, names):
        out = []
        for n in names:
            assert n in self._body_names, f"The specified body name ({n}) doesn't exist: {self._body_names}"
            out.append(self._body_names.index(n))
        return out

    # -- sampling ----------------------------------------------------------------------------------------------------
    def sample_times(self, num_samples, start=False):
        """Reference ``:321-329`` -- numpy GLOBAL RNG, ids drawn first, then the uniform phase."""
        ids = np.random.randint(0, self.num_trajectories, size=num_samples)
        if start:
            return ids, np.zeros(num_samples)
        return ids, np.random.uniform(low=0.0, high=1.0, size=num_samples) * self.durations[ids]

    def compute_frame_blend(self, times, motion_ids):
        return frame_blend_f64(times, motion_ids, self.durations, self.traj_starts, self.traj_ends, self.dt)

    def sample(self, num_samples, times=None, duration=None, motion_ids=None):
        """Reference ``sample`` (``:361-390``); ``duration`` is accepted and unused, as in the reference."""
        if times is None:
            drawn_ids, times = self.sample_times(num_samples)
            if motion_ids is None:
                motion_ids = drawn_ids
        elif motion_ids is None:
            motion_ids = np.zeros(num_samples, dtype=np.int32)
        i0, i1, blend = self.compute_frame_blend(times, motion_ids)
        w = torch.tensor(blend, dtype=torch.float32, device=self.device)

        def _lin(t):
            return lerp_f32(t[i0], t[i1], w)

        return (
            _lin(self.dof_positions),
            _lin(self.dof_velocities),
            _lin(self.body_positions),
            slerp_f32(self.body_rotations[i0], self.body_rotations[i1], w),
            _lin(self.body_linear_velocities),
            _lin(self.body_angular_velocities),
        )
