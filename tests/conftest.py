import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def synthetic_stack(norb, ntrain, seed, layout):
    """Same recipe as tests/golden/make_golden.py::synthetic_stack."""
    rng = np.random.default_rng(seed)
    b = rng.standard_normal((ntrain, ntrain))
    ovlp = np.eye(ntrain) + 0.01 * (b + b.T)
    one = rng.standard_normal((ntrain, ntrain, norb, norb))
    one = one + one.transpose(1, 0, 2, 3)
    two = rng.standard_normal((ntrain, ntrain, norb * norb, norb * norb)) / norb
    two = two + two.transpose(1, 0, 2, 3)
    two = two + two.transpose(0, 1, 3, 2)
    full = two.reshape((ntrain, ntrain) + (norb,) * 4)
    il = np.tril_indices(ntrain)
    if layout == 6:
        return ovlp, one, full
    if layout == 5:
        return ovlp, one, full[il]
    ic = np.tril_indices(norb * norb)
    comp = two[:, :, ic[0], ic[1]]
    if layout == 3:
        return ovlp, one, comp
    assert layout == 2
    return ovlp, one, comp[il]


def random_civec(na, nb, seed, symmetric=False):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((na, nb))
    if symmetric and na == nb:
        c = c + c.T
    return c / np.linalg.norm(c)


PREDICT_CASES = [(4, 2, 3), (6, 6, 3), (7, 3, 5), (10, 10, 6)]


def load_predict_golden(norb, natm, ntrain):
    return np.load(os.path.join(GOLDEN, f"predict_n{norb}_a{natm}_N{ntrain}.npz"))
