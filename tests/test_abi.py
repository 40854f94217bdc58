"""CPU: the C-ABI library loads and exports every symbol include/yolo_b200.h declares
(no compute calls - there is no GPU here), and the ctypes structs match the header."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "yolo_b200.h")
LIB = os.path.join(ROOT, "tf2_yolo_b200", "libyolo_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("yb_loss_fwd_bwd", "yb_loss_v1_fwd_bwd", "yb_loss_v2_fwd_bwd", "yb_loss_v3_fwd_bwd",
                 "yb_loss_v4_fwd_bwd", "yb_decode", "yb_nms", "yb_pairwise_iou", "yb_kmeans_assign",
                 "yb_map_match", "yb_map_accumulate", "yb_pr_curve", "yb_status_string"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_binding_covers_the_header(lib):
    from tf2_yolo_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()
    assert _native.lib.yb_abi_version() == 2
    assert _native.status_string(0) == "ok"
    assert "workspace" in _native.status_string(-4)
    assert _native.lib.yb_loss_workspace_bytes(3) > 0


def test_struct_layout_matches_the_header(tmp_path):
    """sizeof/offsetof from a C compile of the header == ctypes."""
    from tf2_yolo_b200 import _native
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "yolo_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(yb_loss_params), offsetof(yb_loss_params, inv_batch),'
                   ' sizeof(yb_loss_scale), offsetof(yb_loss_scale, p), sizeof(yb_decode_params),'
                   ' offsetof(yb_decode_params, threshold)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_native.LossParams), _native.LossParams.inv_batch.offset,
            ctypes.sizeof(_native.LossScale), _native.LossScale.p.offset,
            ctypes.sizeof(_native.DecodeParams), _native.DecodeParams.threshold.offset]
    assert got == want


def test_host_side_argument_validation_needs_no_gpu(lib):
    """Invalid arguments are rejected on the host before any CUDA call."""
    from tf2_yolo_b200 import _native as N
    assert N.lib.yb_loss_fwd_bwd(None, 1, None, None, None, 0, None) == -1          # YB_E_NULL
    assert N.lib.yb_nms(None, None, 0, 0, 0, 0.5, 1, None, None, None, None, None, 0, None) < 0
    assert N.lib.yb_kmeans_workspace_bytes(10, 0, 2) == 0
    assert N.lib.yb_kmeans_assign(None, 10, 2, None, 9, 0, None, None, None, None, 0, None) == -1
    p = N.DecodeParams()
    p.version, p.class_num, p.n_scales = 9, 80, 1
    assert N.lib.yb_decode((ctypes.c_void_p * 1)(1), 1, ctypes.byref(p), None, 0, ctypes.c_void_p(8), None, 0, None) == -3
