"""In-tree build of ``libamp_b200.so`` (hand-written CUDA for sm_100a behind a C ABI, see ``include/amp_b200.h``).

    python -m humanoid_amp_b200.build            # incremental
    python -m humanoid_amp_b200.build --force

nvcc cross-compiles for sm_100a without a GPU.  The shared object is written next to this file so it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).  The CUDA runtime is linked statically so the
library loads (and its symbols can be checked) on a host without any CUDA driver.
"""

from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libamp_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# per-file extra flags: the motion/obs kernels must not contract a*b+c into FMA (reference rounds every op separately)
SOURCES = {
    "amp_core.cu": [],
    "amp_memory.cu": [],
    # AMP_COLLECT_PROFILE=1 (developer builds only) adds the phase-skipping modes of the reference-motion kernel
    "amp_motion.cu": ["-fmad=false"] + (["-DAMP_COLLECT_PROFILE"] if os.environ.get("AMP_COLLECT_PROFILE") == "1" else []),
    # AMP_DISC_PROFILE=1 (developer builds only) adds in-kernel cycle counters to the fused discriminator kernel
    "amp_disc.cu": ["-DAMP_DISC_PROFILE"] if os.environ.get("AMP_DISC_PROFILE") == "1" else [],
    # AMP_EXCHANGE_STAMPS=1|2 (developer builds only) moves the fused exchange kernel's time stamps inside its push / reduce phase
    "amp_disc_train.cu": ["-DAMP_EXCHANGE_STAMPS=" + os.environ["AMP_EXCHANGE_STAMPS"]] if os.environ.get("AMP_EXCHANGE_STAMPS") else [],
    "amp_bucket.cu": [],
    # the float32 steps of the dataset tool mirror numpy: no contraction into FMA
    "amp_dataset.cu": ["-fmad=false"],
}
HEADERS = ["amp_internal.h", "amp_math.cuh", os.path.join(ROOT, "include", "amp_b200.h")]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def _digest(paths, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    headers += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h")) and os.path.join(CSRC, f) not in headers]
    objects, rebuilt = [], False
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        want = _digest([path, *headers], " ".join(ARCH + COMMON + extra))
        have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(obj) else ""
        if force or want != have:
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", path, "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
            with open(stamp, "w") as f:
                f.write(want)
            rebuilt = True
        objects.append(obj)
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objects]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
