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


TFSTUB_SO = os.path.join(ROOT, "build", "tf_stub", "libyolo_b200_tfstub.so")


def build_tfstub():
    """tf_ops/yolo_loss_op.cc compiled against tests/tf_stub/ (a functional stand-in for the slice of
    the TensorFlow op API it uses) and linked with libyolo_b200.so; returns the ctypes library."""
    import ctypes
    import subprocess
    srcs = [os.path.join(ROOT, "tf2_yolo_b200", "tf_ops", "yolo_loss_op.cc"),
            os.path.join(ROOT, "tests", "tf_stub", "harness.cc")]
    deps = srcs + [os.path.join(ROOT, "tests", "tf_stub", "tf_stub.h"), os.path.join(ROOT, "include", "yolo_b200.h")]
    libdir = os.path.join(ROOT, "tf2_yolo_b200")
    if not os.path.exists(os.path.join(libdir, "libyolo_b200.so")):
        import __graft_entry__ as g
        g.build()
    if (not os.path.exists(TFSTUB_SO)) or any(os.path.getmtime(d) > os.path.getmtime(TFSTUB_SO) for d in deps):
        os.makedirs(os.path.dirname(TFSTUB_SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-comment",
                               "-I", os.path.join(ROOT, "tests", "tf_stub"), "-I", os.path.join(ROOT, "include"),
                               *srcs, "-o", TFSTUB_SO, "-L", libdir, "-l:libyolo_b200.so", f"-Wl,-rpath,{libdir}"])
    lib = ctypes.CDLL(TFSTUB_SO)
    lib.tfstub_last_error.restype = ctypes.c_char_p
    lib.tfstub_describe_ops.restype = ctypes.c_char_p
    lib.tfstub_construct.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    lib.tfstub_run.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                               ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    return lib


def attr_spec(attrs):
    """Attribute dict (as the Python binding hands it to the op) -> harness spec string."""
    parts = []
    for k, v in attrs.items():
        if isinstance(v, bool):
            parts.append(f"{k}=b:{int(v)}")
        elif isinstance(v, int):
            parts.append(f"{k}=i:{v}")
        elif isinstance(v, float):
            parts.append(f"{k}=f:{v!r}")
        elif isinstance(v, (list, tuple)):
            kind = "li" if (len(v) and all(isinstance(x, int) and not isinstance(x, bool) for x in v)) else "lf"
            if k in ("grid_h", "grid_w"):
                kind = "li"
            parts.append(f"{k}={kind}:" + ",".join(repr(x) for x in v))
        else:
            raise TypeError((k, v))
    return ";".join(parts).encode()
