import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_masks(npz, step=0):
    out = {}
    for key in ("embed", "attn", "fc1"):
        shape = tuple(int(x) for x in npz[f"mask{step}_{key}_shape"])
        n = int(np.prod(shape))
        bits = np.unpackbits(npz[f"mask{step}_{key}"])[:n].reshape(shape)
        out[key] = torch.from_numpy(bits.astype(np.bool_))
    return out


@pytest.fixture(scope="session")
def golden_small():
    return load_npz("ref_small.npz")


@pytest.fixture(scope="session")
def golden_default():
    return load_npz("ref_default.npz")


def state_from_npz(npz, prefix):
    from oracle import afr_oracle as orc
    return {k: torch.from_numpy(npz[f"{prefix}/{k}"].copy()) for k in orc.STATE_KEYS}


def rel_fro(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
