"""Shared test plumbing.

Markers
-------
``gpu``  — needs a CUDA device (the parity tests proper; they call the CUDA operators through
           the C-ABI and compare with the CPU oracle / golden fixtures).  Everything else runs
           on the CPU: oracle vs golden fixtures, host logic, ABI symbol checks, gloo tests.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

# PeerGroup.virtual runs up to 8 ranks as 8 streams of ONE device whose kernels wait for each other: give every
# stream its own hardware queue (the default of 8 connections lets unrelated streams alias and serialise).
# Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """golden('adain') → dict of numpy arrays from tests/golden/adain.npz"""
    cache = {}

    def load(name: str):
        if name not in cache:
            with np.load(GOLDEN_DIR / f"{name}.npz") as z:
                cache[name] = {k: z[k] for k in z.files}
        return cache[name]

    return load


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda", 0)


def assert_close_scaled(actual, expected, rtol, name=""):
    """|a-b| <= rtol * max(|expected|, scale) elementwise, scale = mean |expected| of the tensor:
    the 1e-5-relative bar of the north star, made meaningful for values that cross zero
    (SURVEY.md §7 'Variance numerics')."""
    a = torch.as_tensor(np.asarray(actual) if not torch.is_tensor(actual) else actual).double().cpu()
    e = torch.as_tensor(np.asarray(expected) if not torch.is_tensor(expected) else expected).double().cpu()
    assert a.shape == e.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    nan_a, nan_e = torch.isnan(a), torch.isnan(e)
    assert torch.equal(nan_a, nan_e), f"{name}: NaN pattern differs"
    a, e = a[~nan_a], e[~nan_e]
    if e.numel() == 0:
        return
    scale = e.abs().mean().clamp_min(1e-30)
    bound = rtol * torch.maximum(e.abs(), scale)
    err = (a - e).abs()
    bad = err > bound
    assert not bad.any(), (f"{name}: {int(bad.sum())}/{e.numel()} elements exceed rtol={rtol}: "
                           f"max err {err.max().item():.3e}, max ratio {(err / bound).max().item():.2f}")
