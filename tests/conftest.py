import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def loss_cases(npz):
    for name in npz["names"]:
        name = str(name)
        meta = json.loads(str(npz[name + "/meta"]))
        kw = dict(meta["kwargs"])
        is_arr = kw.pop("binary_weight_is_array", False)
        if is_arr:
            kw["binary_weight"] = np.asarray(kw["binary_weight"], dtype=np.float64)
        yield name, meta, kw, npz[name + "/y_true"], npz[name + "/y_pred"], npz[name + "/loss"], npz[name + "/grad"]
