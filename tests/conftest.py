"""Shared pytest plumbing: the ``gpu`` marker, fixture loaders, and skip logic for a box without a B200."""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CLIP_NAMES = [
    "G1_walk",
    "G1_dance",
    "G1_dance_old",
    "G1_walk_lafan1",
    "custom_motion",
    "humanoid_walk",
    "humanoid_run",
    "humanoid_dance",
]
REFERENCE_ROOT = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def clip_path(name: str) -> str:
    return os.path.join(GOLDEN, "clips", f"{name}.npz")


def pooled_spec() -> str:
    return ",".join(clip_path(n) for n in ("humanoid_walk", "humanoid_run", "humanoid_dance"))


def full_clip_path(name: str) -> str:
    """The shipped clip itself (``tests/golden/clips_full``: data files of the reference, committed as fixtures)."""
    return os.path.join(GOLDEN, "clips_full", f"{name}.npz")


def full_pooled_spec() -> str:
    return ",".join(full_clip_path(n) for n in ("humanoid_walk", "humanoid_run", "humanoid_dance"))


def fixture_files(tag: str) -> list:
    """Clip files behind a fixture tag of ``vectors.npz``: ``<name>`` (40-frame window), ``full/<name>`` (whole clip),
    ``pooled_humanoid`` / ``full/pooled_humanoid`` (three clips as one comma-separated spec)."""
    full = tag.startswith("full/")
    name = tag[5:] if full else tag
    if name == "pooled_humanoid":
        return (full_pooled_spec() if full else pooled_spec()).split(",")
    return [full_clip_path(name) if full else clip_path(name)]


FIXTURE_TAGS = CLIP_NAMES + ["pooled_humanoid"] + [f"full/{n}" for n in CLIP_NAMES] + ["full/pooled_humanoid"]


@pytest.fixture(scope="session")
def golden():
    with np.load(os.path.join(GOLDEN, "vectors.npz")) as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)
