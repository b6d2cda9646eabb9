"""ctypes binding of ``libamp_b200.so`` (C ABI declared in ``include/amp_b200.h``).

There is no CPU fallback and no alternative backend: if the shared object is missing, or a call fails, this module
raises.  The library is built in-tree by ``python -m humanoid_amp_b200.build`` (``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libamp_b200.so")

AMP_OK, AMP_EINVAL, AMP_ECUDA, AMP_ENODEV, AMP_ERANGE, AMP_ENOMEM = 0, -1, -2, -3, -4, -5
ABI_VERSION = 1


class AmpB200Error(RuntimeError):
    """A call into libamp_b200.so failed (``code`` is the negative AMP_E* status)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[amp_b200 {code}] {message}")
        self.code = code


class LibDesc(C.Structure):
    """``amp_lib_desc_t``"""

    _fields_ = [
        ("num_frames", C.c_int64),
        ("num_dofs", C.c_int32),
        ("num_bodies", C.c_int32),
        ("num_trajectories", C.c_int32),
        ("_pad0", C.c_int32),
        ("dt", C.c_double),
        ("traj_starts", C.c_void_p),
        ("traj_ends", C.c_void_p),
        ("durations", C.c_void_p),
        ("dof_positions", C.c_void_p),
        ("dof_velocities", C.c_void_p),
        ("body_positions", C.c_void_p),
        ("body_rotations", C.c_void_p),
        ("body_linear_velocities", C.c_void_p),
        ("body_angular_velocities", C.c_void_p),
        ("dof_indexes", C.c_void_p),
        ("num_obs_dofs", C.c_int32),
        ("ref_body_index", C.c_int32),
        ("key_body_indexes", C.c_void_p),
        ("num_key_bodies", C.c_int32),
        ("_pad1", C.c_int32),
    ]


class EnvStepArgs(C.Structure):
    """``amp_env_step_t``"""

    _fields_ = [
        ("joint_pos", C.c_void_p), ("joint_vel", C.c_void_p),
        ("body_pos_w", C.c_void_p), ("body_quat_w", C.c_void_p), ("body_lin_vel_w", C.c_void_p), ("body_ang_vel_w", C.c_void_p),
        ("num_envs", C.c_int64),
        ("num_dofs", C.c_int32), ("num_sim_bodies", C.c_int32), ("ref_body", C.c_int32), ("num_key_bodies", C.c_int32),
        ("key_bodies", C.c_void_p),
        ("num_amp_observations", C.c_int32), ("_pad0", C.c_int32),
        ("amp_buf", C.c_void_p),
        ("last_actions", C.c_void_p), ("command", C.c_void_p),
        ("action_size", C.c_int32), ("command_size", C.c_int32), ("num_actor_observations", C.c_int32),
        ("hist_include_actions", C.c_int32), ("hist_include_command", C.c_int32), ("_pad1", C.c_int32),
        ("hist_buf", C.c_void_p), ("just_reset", C.c_void_p), ("actor_obs", C.c_void_p),
        ("actor_stride", C.c_int64),
        ("reward_scales", C.c_void_p), ("reset_terminated", C.c_void_p), ("actions", C.c_void_p), ("soft_limits", C.c_void_p),
        ("joint_acc", C.c_void_p),
        ("reward_total", C.c_void_p), ("reward_terms", C.c_void_p), ("track_err", C.c_void_p),
    ]  # fmt: skip


class DatasetDesc(C.Structure):
    """``amp_dataset_desc_t`` of include/amp_b200.h."""

    _fields_ = [
        ("n_in", C.c_int32), ("n_cols", C.c_int32), ("n_out", C.c_int32), ("n_dofs", C.c_int32), ("n_bodies", C.c_int32), ("n_joints", C.c_int32),
        ("rows", C.c_void_p), ("t_orig", C.c_void_p), ("t_new", C.c_void_p), ("lerp_lo", C.c_void_p), ("slerp_ind", C.c_void_p),
        ("slerp_alpha", C.c_void_p), ("joint_parent", C.c_void_p), ("joint_qidx", C.c_void_p), ("joint_origin_xyz", C.c_void_p),
        ("joint_origin_rot", C.c_void_p), ("joint_axis", C.c_void_p), ("body_joint", C.c_void_p),
    ]  # fmt: skip


_P, _I32, _I64, _F32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/amp_b200.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "amp_b200_abi_version": (C.c_int, []),
    "amp_last_error": (C.c_char_p, []),
    "amp_set_device": (C.c_int, [C.c_int]),
    "amp_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "amp_lib_create": (C.c_int, [C.POINTER(LibDesc), _P, C.POINTER(_P)]),
    "amp_lib_destroy": (C.c_int, [_P]),
    "amp_lib_obs_width": (C.c_int, [_P]),
    "amp_lib_poll_flags": (C.c_int, [_P, _P, C.POINTER(C.c_uint32)]),
    "amp_lib_set_option": (C.c_int, [_P, _I32, _I64]),
    "amp_frame_blend": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "amp_sample_full": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "amp_lerp": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P]),
    "amp_slerp": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P]),
    "amp_collect_reference": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _I64, _I64, _I64, _P, _P]),
    "amp_compute_obs": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P]),
    "amp_tangent_normal": (C.c_int, [_P, _I64, _P, _P]),
    "amp_obs_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _I32, _I32, _P, _P, _I64, _P]),
    "amp_actor_obs_step": (C.c_int, [_P, _I64, _I32, _I32, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _I64, _P]),
    "amp_task_reward": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _I32, _I32, _P, _I64, _P, _P, _P, _P]),
    "amp_env_step": (C.c_int, [C.POINTER(EnvStepArgs), _P]),
    "amp_disc_create": (C.c_int, [_I32, _I32, _I32, _I64, _P, C.POINTER(_P)]),
    "amp_disc_destroy": (C.c_int, [_P]),
    "amp_disc_chunk_rows": (C.c_int64, [_P]),
    "amp_disc_launch_count": (C.c_int64, [_P, _I64]),
    "amp_disc_load": (C.c_int, [_P] * 9 + [_P]),
    "amp_disc_style_reward": (C.c_int, [_P, _P, _I64, _I64, _F32, _P, _P, _P]),
    "amp_disc_style_reward_indexed": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _F32, _P, _P, _P, _P]),
    "amp_style_reward_from_logits": (C.c_int, [_P, _I64, _F32, _P, _P]),
    "amp_gather_rows": (C.c_int, [_P, _I64, _I64, _P, _I64, _I32, _P, _I64, _P, _P]),
    "amp_scaler_scratch_bytes": (C.c_int64, [_I32]),
    "amp_scaler_update": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _P, _I64, _P]),
    "amp_scaler_apply": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _F32, _F32, _P, _I64, _P]),
    "amp_disc_train_create": (C.c_int, [_I32, _I32, _I32, _I64, _P, C.POINTER(_P)]),
    "amp_disc_train_destroy": (C.c_int, [_P]),
    "amp_disc_train_stage": (C.c_int, [_P, _I32, _P, _I64, _I64, _P, _P, _P]),
    "amp_disc_train_step": (C.c_int, [_P] * 7 + [_I64, _F32, _F32, _F32, _F32] + [_P] * 9),
    "amp_disc_train_step_exchange": (C.c_int, [_P] * 7 + [_I64, _F32, _F32, _F32, _F32] + [_P] * 10),
    "amp_dataset_scratch_bytes": (C.c_int64, [_I32, _I32, _I32]),
    "amp_dataset_interp_fk": (C.c_int, [C.POINTER(DatasetDesc), _P, _P, _P, _P, _P, _I64, _P]),
    "amp_dataset_velocities": (C.c_int, [_I32, _I32, _I32, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "amp_bucket_create": (C.c_int, [_I64, _I32, _I32, C.POINTER(_P)]),
    "amp_bucket_create_shared": (C.c_int, [_I64, _I32, _I32, C.POINTER(_P)]),
    "amp_bucket_export_shared": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "amp_bucket_join_shared": (C.c_int, [_P, _I32]),
    "amp_bucket_connect_shared": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "amp_bucket_in_switch": (C.c_int, [_P]),
    "amp_bucket_destroy": (C.c_int, [_P]),
    "amp_bucket_floats": (C.c_int64, [_P]),
    "amp_bucket_data": (C.c_void_p, [_P]),
    "amp_bucket_export": (C.c_int, [_P, _P]),
    "amp_bucket_connect": (C.c_int, [_P, _P]),
    "amp_bucket_allreduce_mean": (C.c_int, [_P, _I64, _I64, _P]),
    "amp_bucket_poll_status": (C.c_int, [_P, _P, C.POINTER(C.c_uint32)]),
    "amp_bucket_last_timing": (C.c_int, [_P, _P, C.POINTER(C.c_uint64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AmpB200Error(
            AMP_ENODEV,
            f"{LIB_PATH} is missing: build the CUDA extension with `python -m humanoid_amp_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)",
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale relative to the header
        fn.restype, fn.argtypes = res, args
    got = lib.amp_b200_abi_version()
    if got != ABI_VERSION:
        raise AmpB200Error(AMP_EINVAL, f"libamp_b200.so has ABI version {got}, the Python shim expects {ABI_VERSION}")
    _lib = lib
    return lib


_restore = threading.local()  # device to hand back to the calling thread after the library call that enter() prepared


def _restore_device() -> None:
    prev = getattr(_restore, "device", None)
    if prev is not None:
        _restore.device = None
        load().amp_set_device(prev)


def check(status: int) -> None:
    """Raise on a failed library call; also puts the thread's current CUDA device back where it was before :func:`enter`
    (the library's runtime shares the driver's per-thread current context with torch's, so a call on another device would
    otherwise silently change ``torch.cuda.current_device()``)."""
    msg = load().amp_last_error() if status != AMP_OK else None
    _restore_device()
    if status != AMP_OK:
        raise AmpB200Error(status, msg.decode(errors="replace") if msg else "unknown error")


def require_cuda(device) -> torch.device:
    """The product path runs on a CUDA device only."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise AmpB200Error(AMP_ENODEV, f"humanoid_amp_b200 runs on CUDA devices only (got device={device!r}); there is no CPU path")
    if not torch.cuda.is_available():
        raise AmpB200Error(AMP_ENODEV, "no CUDA device is available; humanoid_amp_b200 has no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def enter(device: torch.device):
    """Make ``device`` current in the library's runtime and return (lib, cudaStream_t of torch's current stream)."""
    lib = load()
    _restore_device()  # a previous enter() whose call was never check()ed
    current = torch.cuda.current_device()
    if current != device.index:
        status = lib.amp_set_device(device.index)
        if status != AMP_OK:
            check(status)
        _restore.device = current
    return lib, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a tensor / numpy array, or NULL."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)
