"""``MotionLoader`` with the reference's call surface, executing on hand-written sm_100a kernels.

Mirrors ``/root/reference/motions/motion_loader.py`` (class ``:87-430``, file resolution ``:14-84``): same constructor,
attributes, method names, argument meaning, return types and error behaviour, so the reference's callers
(``g1_amp_env.py:35-60, 371-486``) can switch to it unchanged.  The arithmetic runs in ``libamp_b200.so``
(``csrc/amp_motion.cu``); there is no CPU path.

Differences that are deliberate and documented:

* ``times`` / ``motion_ids`` may also be CUDA tensors (float64 / int64) -- then nothing crosses PCIe;
* out-of-range ids in a *device* tensor cannot raise ``IndexError`` synchronously; the kernels clamp them and raise a
  sticky flag readable with :meth:`MotionLoader.poll_flags`.  Host arrays are validated like numpy indexing would.
"""

from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Optional, Sequence

import numpy as np
import torch
import yaml

from . import _lib
from .clip_cache import TENSOR_KEYS as _TENSOR_KEYS
from .clip_cache import load_clips


def _resolve_motion_files(motion_file: str) -> list[str]:
    """Turn a motion spec into a file list, with the reference's precedence (``motion_loader.py:14-84``):

    YAML config (``motion_files`` list, else ``glob_pattern``; relative entries resolve against the config's directory)
    -> comma separated paths -> glob pattern -> directory of ``*.npz`` -> single file.  ``ValueError`` if nothing matches.
    """
    if motion_file.endswith((".yaml", ".yml")):
        base = os.path.dirname(motion_file)
        with open(motion_file, "r") as fh:
            cfg = yaml.safe_load(fh)
        found: list[str] = []
        for entry in (cfg or {}).get("motion_files", None) or []:
            full = entry if os.path.isabs(entry) else os.path.join(base, entry)
            if os.path.exists(full):
                found.append(full)
            else:
                print(f"Warning: File not found: {full}")
        if not found and cfg and "glob_pattern" in cfg:
            pattern = cfg["glob_pattern"]
            found = sorted(glob.glob(pattern if os.path.isabs(pattern) else os.path.join(base, pattern)))
        if not found:
            raise ValueError(f"No valid motion files found in config: {motion_file}")
        return found

    if "," in motion_file:
        listed = [p.strip() for p in motion_file.split(",")]
        listed = [p for p in listed if os.path.exists(p)]
        if listed:
            return listed
    if "*" in motion_file or "?" in motion_file:
        matched = sorted(glob.glob(motion_file))
        if matched:
            return matched
    if os.path.isdir(motion_file):
        inside = sorted(glob.glob(os.path.join(motion_file, "*.npz")))
        if inside:
            return inside
    if os.path.exists(motion_file):
        return [motion_file]
    raise ValueError(f"No files found for pattern: {motion_file}")


class _LibHandle:
    """Owns one ``amp_lib_t`` (a staged motion library, optionally with the env's column selection)."""

    def __init__(self, loader: "MotionLoader", dof_indexes=None, ref_body_index=None, key_body_indexes=None):
        self.device = loader.device
        lib, stream = _lib.enter(self.device)
        # host arrays must outlive the create call only
        starts = np.ascontiguousarray(loader.traj_starts, dtype=np.int64)
        ends = np.ascontiguousarray(loader.traj_ends, dtype=np.int64)
        durs = np.ascontiguousarray(loader.durations, dtype=np.float64)
        d = _lib.LibDesc()
        d.num_frames = loader.num_frames
        d.num_dofs = loader.num_dofs
        d.num_bodies = loader.num_bodies
        d.num_trajectories = loader.num_trajectories
        d.dt = float(loader.dt)
        d.traj_starts, d.traj_ends, d.durations = starts.ctypes.data, ends.ctypes.data, durs.ctypes.data
        for key in _TENSOR_KEYS:
            setattr(d, key, getattr(loader, key).data_ptr())
        keep = [starts, ends, durs]
        if dof_indexes is not None:
            dof = np.ascontiguousarray(dof_indexes, dtype=np.int32)
            keys = np.ascontiguousarray(key_body_indexes, dtype=np.int32)
            keep += [dof, keys]
            d.dof_indexes, d.num_obs_dofs = dof.ctypes.data, len(dof)
            d.ref_body_index = int(ref_body_index)
            d.key_body_indexes, d.num_key_bodies = (keys.ctypes.data if len(keys) else None), len(keys)
        handle = C.c_void_p()
        _lib.check(lib.amp_lib_create(C.byref(d), stream, C.byref(handle)))
        self._h = handle
        self._loader = loader  # keeps the six clip tensors (borrowed by the handle) alive
        self.obs_width = lib.amp_lib_obs_width(handle)

    @property
    def raw(self) -> C.c_void_p:
        if self._h is None:
            raise _lib.AmpB200Error(_lib.AMP_EINVAL, "motion library handle was destroyed")
        return self._h

    COLLECT_TABLE = {"auto": 0, "global": 1, "smem": 2}

    def set_collect_table(self, mode: str) -> None:
        """``AMP_OPT_COLLECT_TABLE``: where the fused collect kernel keeps the packed row table (``auto`` | ``global`` |
        ``smem``).  A tuning / test knob: results do not depend on it."""
        lib, _ = _lib.enter(self.device)
        _lib.check(lib.amp_lib_set_option(self.raw, 1, self.COLLECT_TABLE[mode]))

    def close(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().amp_lib_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone; the driver reclaims the handle
            pass


class MotionLoader:
    """Load motion clips in the reference ``.npz`` format and sample them on the GPU.

    Reference: ``motions/motion_loader.py:87-430``.  ``device`` must be a CUDA device.
    """

    def __init__(self, motion_file: str, device, cache_dir: Optional[str] = None) -> None:
        """``motion_file``: anything the reference accepts (``motion_loader.py:14-84``) or one packed ``.ampclip`` file.
        ``cache_dir`` (or the environment variable ``AMP_B200_CLIP_CACHE``): keep / reuse a packed cache of the resolved
        clip set there (``clip_cache.py``)."""
        files = _resolve_motion_files(motion_file)
        print(f"Loading {len(files)} motion file(s) from: {motion_file}")
        self.device = _lib.require_cuda(device)

        def pinned(nbytes: int) -> np.ndarray:
            self._staging = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            return self._staging.numpy()

        clip, self.load_path = load_clips(files, cache_dir or os.environ.get("AMP_B200_CLIP_CACHE") or None, pinned)
        self._dof_names, self._body_names = clip.dof_names, clip.body_names
        self.dt = clip.dt
        self.traj_starts, self.traj_ends, self.durations = clip.traj_starts, clip.traj_ends, clip.durations
        self.num_trajectories = clip.num_trajectories
        self.num_frames = clip.num_frames
        self.duration = float(np.sum(self.durations))
        # ONE host-to-device copy of the whole arena; the six tensors the reference keeps (:141-158) are views into it
        self._arena = self._staging.to(self.device, non_blocking=True)
        for k, off, shape in zip(_TENSOR_KEYS, clip.offsets, clip.shapes):
            nbytes = 4 * int(np.prod(shape))
            setattr(self, k, self._arena[off : off + nbytes].view(torch.float32).view(shape))
        print(
            f"Motion loaded: {self.num_trajectories} files, total duration: {self.duration} sec, total frames: {self.num_frames}"
        )
        self._plain = _LibHandle(self)

    # ---- metadata (reference :166-184, :392-430) -------------------------------------------------------------------
    @property
    def dof_names(self) -> list[str]:
        return self._dof_names

    @property
    def body_names(self) -> list[str]:
        return self._body_names

    @property
    def num_dofs(self) -> int:
        return len(self._dof_names)

    @property
    def num_bodies(self) -> int:
        return len(self._body_names)

    def get_dof_index(self, dof_names: Sequence[str]) -> list[int]:
        out = []
        for name in dof_names:
            assert name in self._dof_names, f"The specified DOF name ({name}) doesn't exist: {self._dof_names}"
            out.append(self._dof_names.index(name))
        return out

    def get_body_index(self, body_names: Sequence[str]) -> list[int]:
        out = []
        for name in body_names:
            assert name in self._body_names, f"The specified body name ({name}) doesn't exist: {self._body_names}"
            out.append(self._body_names.index(name))
        return out

    # ---- argument marshalling --------------------------------------------------------------------------------------
    def _times_to_device(self, times) -> torch.Tensor:
        if isinstance(times, torch.Tensor):
            if times.device != self.device or times.dtype != torch.float64:
                # a pinned host tensor is copied asynchronously on the current stream (the caller keeps it alive, as with
                # any non_blocking copy); pageable memory falls back to torch's staged synchronous copy
                times = times.to(device=self.device, dtype=torch.float64, non_blocking=times.device.type == "cpu" and times.is_pinned())
            return times.contiguous().view(-1)
        host = np.ascontiguousarray(np.asarray(times, dtype=np.float64).reshape(-1))
        return torch.from_numpy(host).to(self.device, non_blocking=False)

    def _ids_to_device(self, motion_ids, count: int) -> Optional[torch.Tensor]:
        """``None`` means all-zero ids (reference ``:366``).  Host ids get numpy's indexing semantics."""
        if motion_ids is None:
            return None
        if isinstance(motion_ids, torch.Tensor):
            pinned = motion_ids.device.type == "cpu" and motion_ids.is_pinned()
            ids = motion_ids.to(device=self.device, dtype=torch.int64, non_blocking=pinned).contiguous().view(-1)
        else:
            host = np.asarray(motion_ids).reshape(-1)
            if host.size and not np.issubdtype(host.dtype, np.integer):
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            host = host.astype(np.int64)
            n = self.num_trajectories
            if host.size and (host.min() < -n or host.max() >= n):
                bad = host[(host < -n) | (host >= n)][0]
                raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n}")
            host = np.where(host < 0, host + n, host)  # numpy wraps negative indices
            ids = torch.from_numpy(np.ascontiguousarray(host)).to(self.device)
        if ids.numel() != count:
            raise ValueError(f"operands could not be broadcast together with shapes ({count},) ({ids.numel()},)")
        return ids

    def poll_flags(self) -> int:
        """Read and clear the device-side sticky flags (bit0: id out of range, bit1: NaN time).  Synchronises."""
        lib, stream = _lib.enter(self.device)
        flags = C.c_uint32(0)
        _lib.check(lib.amp_lib_poll_flags(self._plain.raw, stream, C.byref(flags)))
        return flags.value

    # ---- sampling --------------------------------------------------------------------------------------------------
    def sample_times(self, num_samples: int, start: bool = False) -> tuple[np.ndarray, np.ndarray]:
        """Reference ``:309-329``.  Host numpy GLOBAL RNG on purpose: with the same seed the stream equals the
        reference's (ids are drawn first, then the uniform phase)."""
        motion_ids = np.random.randint(0, self.num_trajectories, size=num_samples)
        if start:
            times = np.zeros(num_samples)
        else:
            times = np.random.uniform(low=0.0, high=1.0, size=num_samples) * self.durations[motion_ids]
        return motion_ids, times

    def sample_times_device(self, num_samples: int, start: bool = False, generator: Optional[torch.Generator] = None):
        """Device-resident variant (Philox): distribution-equal, not stream-equal, to :meth:`sample_times`."""
        ids = torch.randint(0, self.num_trajectories, (num_samples,), device=self.device, generator=generator)
        if start:
            return ids, torch.zeros(num_samples, dtype=torch.float64, device=self.device)
        durs = torch.from_numpy(self.durations).to(self.device)
        u = torch.rand(num_samples, dtype=torch.float64, device=self.device, generator=generator)
        return ids, u * durs[ids]

    def _compute_frame_blend(self, times, motion_ids) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Reference ``:281-307`` evaluated in float64 on the device (bit-exact); returns host arrays like the reference."""
        i0, i1, _, b64 = self.compute_frame_blend_device(times, motion_ids, want_blend64=True)
        return i0.cpu().numpy(), i1.cpu().numpy(), b64.cpu().numpy()

    def compute_frame_blend_device(self, times, motion_ids, want_blend64: bool = False):
        t = self._times_to_device(times)
        n = t.numel()
        ids = self._ids_to_device(motion_ids, n)
        i0 = torch.empty(n, dtype=torch.int64, device=self.device)
        i1 = torch.empty(n, dtype=torch.int64, device=self.device)
        b32 = torch.empty(n, dtype=torch.float32, device=self.device)
        b64 = torch.empty(n, dtype=torch.float64, device=self.device) if want_blend64 else None
        lib, stream = _lib.enter(self.device)
        _lib.check(
            lib.amp_frame_blend(self._plain.raw, _lib.ptr(t), _lib.ptr(ids), n, _lib.ptr(i0), _lib.ptr(i1), _lib.ptr(b32), _lib.ptr(b64), stream)
        )
        return i0, i1, b32, b64

    def sample(
        self,
        num_samples: int,
        times=None,
        duration: float | None = None,
        motion_ids=None,
    ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """Reference ``:331-390``: dof positions, dof velocities, body positions, body rotations (wxyz), body linear and
        angular velocities at the given times, fp32 on the device.  ``duration`` is accepted and unused (as upstream)."""
        if times is None:
            drawn_ids, times = self.sample_times(num_samples)
            if motion_ids is None:
                motion_ids = drawn_ids
        t = self._times_to_device(times)
        n = t.numel()
        ids = self._ids_to_device(motion_ids, n)  # None -> zeros, like np.zeros(num_samples, int32) upstream
        D, B = self.num_dofs, self.num_bodies
        f32 = dict(dtype=torch.float32, device=self.device)
        outs = (
            torch.empty((n, D), **f32),
            torch.empty((n, D), **f32),
            torch.empty((n, B, 3), **f32),
            torch.empty((n, B, 4), **f32),
            torch.empty((n, B, 3), **f32),
            torch.empty((n, B, 3), **f32),
        )
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_sample_full(self._plain.raw, _lib.ptr(t), _lib.ptr(ids), n, *[_lib.ptr(o) for o in outs], stream))
        return outs

    # ---- explicit end-point helpers (reference :186-279) -------------------------------------------------------------
    def _prep_pair(self, a, b, blend, start, end):
        if start is not None and end is not None:
            idx0 = torch.as_tensor(np.asarray(start), device=self.device, dtype=torch.int64)
            idx1 = torch.as_tensor(np.asarray(end), device=self.device, dtype=torch.int64)
            a, b = a.index_select(0, idx0), a.index_select(0, idx1)
        a = a.to(self.device, torch.float32).contiguous()
        b = b.to(self.device, torch.float32).contiguous()
        blend = blend.to(self.device, torch.float32).contiguous().view(-1)
        if a.shape != b.shape or blend.numel() != a.shape[0]:
            raise RuntimeError(f"shape mismatch: a {tuple(a.shape)} b {tuple(b.shape)} blend {tuple(blend.shape)}")
        return a, b, blend

    def _interpolate(self, a, *, b=None, blend=None, start=None, end=None) -> torch.Tensor:
        """``(1.0 - blend) * a + blend * b`` with blend broadcast over trailing dims (reference ``:209-215``)."""
        a, b, blend = self._prep_pair(a, b, blend, start, end)
        out = torch.empty_like(a)
        n = a.shape[0]
        inner = a.numel() // n if n else 1
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_lerp(_lib.ptr(a), _lib.ptr(b), _lib.ptr(blend), n, max(inner, 1), _lib.ptr(out), stream))
        return out

    def _slerp(self, q0, *, q1=None, blend=None, start=None, end=None) -> torch.Tensor:
        """Shortest-arc slerp, wxyz, ``(N,4)`` or ``(N,M,4)`` (reference ``:240-279``)."""
        q0, q1, blend = self._prep_pair(q0, q1, blend, start, end)
        if q0.shape[-1] != 4:
            raise RuntimeError("quaternions must have 4 components in the last dimension")
        out = torch.empty_like(q0)
        n = q0.shape[0]
        bodies = q0.numel() // (4 * n) if n else 1
        lib, stream = _lib.enter(self.device)
        _lib.check(lib.amp_slerp(_lib.ptr(q0), _lib.ptr(q1), _lib.ptr(blend), n, max(bodies, 1), _lib.ptr(out), stream))
        return out

    # ---- used by the env path --------------------------------------------------------------------------------------
    def make_env_handle(self, dof_indexes, ref_body_index, key_body_indexes) -> _LibHandle:
        """Stage the packed AMP rows for the fused ``collect_reference_motions`` kernel."""
        return _LibHandle(self, dof_indexes, ref_body_index, key_body_indexes)
