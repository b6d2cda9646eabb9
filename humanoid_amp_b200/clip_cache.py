"""On-disk motion-clip formats: the reference's ``.npz`` clips and a packed single-file cache.

The reference keeps one ``.npz`` per clip (``motions/README.md:11-21``: ``fps, dof_names, body_names, dof_positions,
dof_velocities, body_positions, body_rotations, body_linear_velocities, body_angular_velocities``) and, on every start,
``MotionLoader.__init__`` (``motions/motion_loader.py:98-164``) unzips each file, concatenates the clips, narrows float64
arrays to float32 and uploads six tensors one by one.  For the 171 k-frame deploy pools that is seconds of host work per
process and per rank.

``.ampclip`` is the result of that work, written once: a small header, the trajectory tables, and the six float32 arrays
back to back in ONE 256-byte-aligned arena, so loading is one ``readinto`` a pinned buffer and ONE host-to-device copy; the
six tensors are views into the device arena.  Byte layout (little endian):

    0    8   magic  b"AMPCLIP1"
    8    4   u32    version (1)
    12   4   u32    header_bytes (offset of the arena, multiple of 256)
    16   8   i64    num_frames            24  4  i32 num_dofs      28  4  i32 num_bodies     32  4  i32 num_trajectories
    36   4   u32    json_bytes            40  8  f64 dt (1 / fps of the FIRST clip, motion_loader.py:122)
    48   8   u64    arena_bytes           56  8  u64 CRC-32 of the arena
    64   6*8 u64    byte offset of each array inside the arena (order of ``TENSOR_KEYS``)
    112  ...        traj_starts i64[T], traj_ends i64[T], durations f64[T], then the JSON blob
                    {"dof_names": [...], "body_names": [...], "sources": [[basename, size, mtime_ns], ...]}

The cache is keyed by the resolved file list (names, sizes, mtimes): a changed or re-ordered source set gets a new cache
file, a stale one is never read.  Everything here is host code (numpy); it needs no GPU and is covered by the CPU tests.
"""

from __future__ import annotations

import hashlib
import json
import os
import struct
import zlib
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

TENSOR_KEYS = (
    "dof_positions",
    "dof_velocities",
    "body_positions",
    "body_rotations",
    "body_linear_velocities",
    "body_angular_velocities",
)
MAGIC = b"AMPCLIP1"
VERSION = 1
_ALIGN = 256
_FIXED = struct.Struct("<8sIIqiiiIdQQ6Q")  # 112 bytes


class ClipCacheError(ValueError):
    """The file is not a valid ``.ampclip`` (bad magic / version / size / checksum)."""


@dataclass
class ClipArrays:
    """What ``MotionLoader.__init__`` leaves behind, on the host: fp32 arrays are views into ``arena``."""

    dof_names: List[str]
    body_names: List[str]
    dt: np.float64
    traj_starts: np.ndarray  # int64 [T]
    traj_ends: np.ndarray  # int64 [T]
    durations: np.ndarray  # float64 [T]
    arena: np.ndarray  # uint8, all six arrays back to back (256-byte aligned offsets)
    offsets: tuple  # byte offset of each array in the arena
    shapes: tuple  # shape of each array

    @property
    def num_frames(self) -> int:
        return int(self.shapes[0][0])

    @property
    def num_trajectories(self) -> int:
        return len(self.traj_starts)

    def array(self, key: str) -> np.ndarray:
        i = TENSOR_KEYS.index(key)
        n = int(np.prod(self.shapes[i]))
        return self.arena[self.offsets[i] : self.offsets[i] + 4 * n].view(np.float32).reshape(self.shapes[i])


def _round_up(n: int, a: int = _ALIGN) -> int:
    return (n + a - 1) // a * a


def _checksum(buf: np.ndarray) -> int:
    """CRC-32 of the arena (zlib, ~GB/s) -- the format's integrity check, stored in a 64-bit field."""
    return zlib.crc32(memoryview(buf)) & 0xFFFFFFFF


def _pack_arena(arrays: Sequence[np.ndarray], pinned_alloc=None):
    shapes = tuple(tuple(int(d) for d in a.shape) for a in arrays)
    offsets, cursor = [], 0
    for a in arrays:
        offsets.append(cursor)
        cursor = _round_up(cursor + a.size * 4)
    arena = pinned_alloc(cursor) if pinned_alloc else np.empty(cursor, dtype=np.uint8)
    arena[:] = 0  # alignment gaps are part of the file and of its checksum
    for a, off in zip(arrays, offsets):
        arena[off : off + a.size * 4].view(np.float32)[:] = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    return arena, tuple(offsets), shapes


def load_npz_clips(files: Sequence[str], pinned_alloc=None) -> ClipArrays:
    """The host side of the reference ``MotionLoader.__init__`` (``motions/motion_loader.py:108-158``): names and ``dt`` come
    from the FIRST file only (``:119-122``), clips are concatenated along the frame axis, ``traj_starts/ends`` are global frame
    indices (``:131-134``), ``durations = dt * (frames - 1)`` (``:135``), and every array is narrowed to float32 the way
    ``torch.tensor(..., dtype=torch.float32)`` does (``:141-158``)."""
    parts = {k: [] for k in TENSOR_KEYS}
    starts, ends, durs = [], [], []
    cursor, dt, dof_names, body_names = 0, None, None, None
    for path in files:
        with np.load(path) as data:
            if dt is None:
                dof_names = data["dof_names"].tolist()
                body_names = data["body_names"].tolist()
                dt = 1.0 / data["fps"]
            for k in TENSOR_KEYS:
                parts[k].append(data[k])
            n_frames = data["dof_positions"].shape[0]
        starts.append(cursor)
        cursor += n_frames
        ends.append(cursor - 1)
        durs.append(dt * (n_frames - 1))
    arrays = [np.concatenate(parts[k]).astype(np.float32) for k in TENSOR_KEYS]
    if arrays[2].shape[1] != len(body_names) or arrays[0].shape[1] != len(dof_names):
        raise ValueError("clip tensors do not match dof_names / body_names of the first file")
    arena, offsets, shapes = _pack_arena(arrays, pinned_alloc)
    return ClipArrays(dof_names, body_names, dt, np.array(starts), np.array(ends), np.array(durs), arena, offsets, shapes)


def _source_fingerprint(files: Sequence[str]) -> list:
    out = []
    for f in files:
        st = os.stat(f)
        out.append([os.path.basename(f), int(st.st_size), int(st.st_mtime_ns)])
    return out


def cache_path_for(files: Sequence[str], cache_dir: str) -> str:
    """Cache file name for a resolved file list: a digest of (absolute path, size, mtime) of every source, in order."""
    h = hashlib.sha256()
    for f in files:
        st = os.stat(f)
        h.update(f"{os.path.abspath(f)}|{st.st_size}|{st.st_mtime_ns}\n".encode())
    stem = os.path.splitext(os.path.basename(files[0]))[0]
    return os.path.join(cache_dir, f"{stem}-{len(files)}clips-{h.hexdigest()[:16]}.ampclip")


def write_clip_cache(path: str, clip: ClipArrays, sources: Optional[Sequence[str]] = None) -> str:
    """Write ``clip`` as a packed ``.ampclip`` file (atomically: temp file + rename)."""
    T = clip.num_trajectories
    blob = json.dumps(
        {"dof_names": list(clip.dof_names), "body_names": list(clip.body_names), "sources": _source_fingerprint(sources or [])}
    ).encode()
    tables = (
        np.ascontiguousarray(clip.traj_starts, dtype="<i8").tobytes()
        + np.ascontiguousarray(clip.traj_ends, dtype="<i8").tobytes()
        + np.ascontiguousarray(clip.durations, dtype="<f8").tobytes()
    )
    header_bytes = _round_up(_FIXED.size + len(tables) + len(blob))
    D, B = clip.shapes[0][1], clip.shapes[2][1]
    fixed = _FIXED.pack(MAGIC, VERSION, header_bytes, clip.num_frames, D, B, T, len(blob), float(clip.dt), clip.arena.size,
                        _checksum(clip.arena), *clip.offsets)
    tmp = f"{path}.tmp{os.getpid()}"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(tmp, "wb") as fh:
        fh.write(fixed + tables + blob)
        fh.write(b"\0" * (header_bytes - _FIXED.size - len(tables) - len(blob)))
        fh.write(memoryview(clip.arena))
    os.replace(tmp, path)
    return path


def read_clip_cache(path: str, pinned_alloc=None, verify: bool = True) -> ClipArrays:
    """Read a ``.ampclip`` file: header + tables with small reads, the arena with ONE ``readinto`` (into pinned memory when
    ``pinned_alloc`` is given).  ``ClipCacheError`` on any inconsistency."""
    size = os.path.getsize(path)
    with open(path, "rb") as fh:
        head = fh.read(_FIXED.size)
        if len(head) < _FIXED.size:
            raise ClipCacheError(f"{path}: truncated header")
        (magic, version, header_bytes, F, D, B, T, json_bytes, dt, arena_bytes, digest, *offsets) = _FIXED.unpack(head)
        if magic != MAGIC:
            raise ClipCacheError(f"{path}: not an .ampclip file (bad magic)")
        if version != VERSION:
            raise ClipCacheError(f"{path}: unsupported version {version}")
        if F < 1 or D < 1 or B < 1 or T < 1 or header_bytes % _ALIGN or header_bytes + arena_bytes != size:
            raise ClipCacheError(f"{path}: inconsistent sizes in header")
        tables = fh.read(24 * T)
        blob = fh.read(json_bytes)
        if len(tables) != 24 * T or len(blob) != json_bytes:
            raise ClipCacheError(f"{path}: truncated tables")
        starts = np.frombuffer(tables, dtype="<i8", count=T, offset=0).astype(np.int64)
        ends = np.frombuffer(tables, dtype="<i8", count=T, offset=8 * T).astype(np.int64)
        durs = np.frombuffer(tables, dtype="<f8", count=T, offset=16 * T).astype(np.float64)
        try:
            meta = json.loads(blob.decode())
            dof_names, body_names = list(meta["dof_names"]), list(meta["body_names"])
        except (ValueError, KeyError) as exc:
            raise ClipCacheError(f"{path}: bad metadata blob") from exc
        shapes = ((F, D), (F, D), (F, B, 3), (F, B, 4), (F, B, 3), (F, B, 3))
        for off, shp in zip(offsets, shapes):
            if off % _ALIGN or off + 4 * int(np.prod(shp)) > arena_bytes:
                raise ClipCacheError(f"{path}: array offset outside the arena")
        if len(dof_names) != D or len(body_names) != B or int(ends[-1]) != F - 1:
            raise ClipCacheError(f"{path}: tables do not match the array shapes")
        arena = pinned_alloc(arena_bytes) if pinned_alloc else np.empty(arena_bytes, dtype=np.uint8)
        fh.seek(header_bytes)
        got = fh.readinto(memoryview(arena))
        if got != arena_bytes:
            raise ClipCacheError(f"{path}: truncated arena")
    if verify and _checksum(arena) != digest:
        raise ClipCacheError(f"{path}: arena checksum mismatch")
    return ClipArrays(dof_names, body_names, np.float64(dt), starts, ends, durs, arena, tuple(offsets), shapes)


def cache_sources(path: str) -> list:
    """The ``[basename, size, mtime_ns]`` records of the clips a cache file was built from."""
    with open(path, "rb") as fh:
        head = fh.read(_FIXED.size)
        (_m, _v, _hb, _F, _D, _B, T, json_bytes, *_rest) = _FIXED.unpack(head)
        fh.seek(_FIXED.size + 24 * T)
        return json.loads(fh.read(json_bytes).decode()).get("sources", [])


def load_clips(files: Sequence[str], cache_dir: Optional[str] = None, pinned_alloc=None) -> tuple:
    """``(ClipArrays, how)``: from the packed cache when ``cache_dir`` holds a valid one for exactly these files, else from
    the ``.npz`` sources (writing the cache when ``cache_dir`` is given).  A single ``.ampclip`` path is read directly."""
    if len(files) == 1 and files[0].endswith(".ampclip"):
        return read_clip_cache(files[0], pinned_alloc), "ampclip"
    if cache_dir:
        path = cache_path_for(files, cache_dir)
        if os.path.exists(path):
            try:
                return read_clip_cache(path, pinned_alloc), "cache-hit"
            except ClipCacheError as exc:
                print(f"Warning: ignoring unusable clip cache ({exc})")
        clip = load_npz_clips(files, pinned_alloc)
        try:
            write_clip_cache(path, clip, files)
        except OSError as exc:
            print(f"Warning: could not write clip cache {path}: {exc}")
        return clip, "cache-miss"
    return load_npz_clips(files, pinned_alloc), "npz"
