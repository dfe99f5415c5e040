"""Shared test plumbing.

Markers
    gpu   needs a B200; run with `pytest -m gpu` on the GPU box.  Everything else runs on CPU.

`-m "not gpu"` covers: the oracle against the golden vectors, the host logic, the boundary headers,
and that the C ABI library loads and exports every declared symbol.  `-m gpu` tests are the parity
tests proper; they call the CUDA path through the C ABI / the C++ class and use the oracle only as
the checker.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_TREE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def netcuda():
    import __graft_entry__

    __graft_entry__.build()
    import netcuda as nc

    return nc


@pytest.fixture(scope="session")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def c1_net(oracle):
    """Config C1: 784-128-64-10 with the reference's random init under srand(1)."""
    npl, n_ins = [128, 64, 10], 784
    w, b = oracle.rand_init(1, 784 * 128 + 128 * 64 + 64 * 10, sum(npl))
    return npl, n_ins, w, b


def rel_err(a, ref):
    """max |a - ref| relative to max |ref| per sample (the north_star's logit tolerance)."""
    a = np.asarray(a, dtype=np.float64).reshape(ref.shape[0], -1)
    r = np.asarray(ref, dtype=np.float64).reshape(ref.shape[0], -1)
    scale = np.maximum(np.abs(r).max(axis=1), 1e-30)
    return float((np.abs(a - r).max(axis=1) / scale).max())
