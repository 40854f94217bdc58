"""GPU: the TensorFlow custom-op kernels (tf_ops/yolo_loss_op.cc), compiled against the TF API
stand-in of tests/tf_stub and run through its harness, give the SAME bits as the ctypes path for
the reference fixtures - i.e. attribute parsing, parameter packing, output allocation and the C-ABI
calls of the op classes are right, for all four packages."""
import ctypes as C
import importlib
import sys

import numpy as np
import pytest
import torch

from conftest import attr_spec, build_tfstub, loss_cases
from test_tf_ops import tf_stub  # noqa: F401  (fixture: recording stand-in for the tensorflow module)
from tf2_yolo_b200.grid_loss import fused_losses

pytestmark = pytest.mark.gpu
PKG = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}


def run_op(lib, op, attrs, inputs, outputs):
    temp = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    ins = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
    elems = (C.c_int64 * len(inputs))(*[t.numel() for t in inputs])
    outs = (C.c_void_p * len(outputs))(*[t.data_ptr() for t in outputs])
    rc = lib.tfstub_run(op.encode(), attr_spec(attrs), len(inputs), ins, elems, len(outputs), outs,
                        C.c_void_p(temp.data_ptr()), temp.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0, lib.tfstub_last_error()


def test_op_kernels_equal_the_ctypes_path_on_the_reference_fixtures(tf_stub, golden):  # noqa: F811
    lib = build_tfstub()
    n = 0
    for name, meta, kw, yt, yp, loss_ref, grad_ref in loss_cases(golden("loss")):
        ver, S = meta["version"], meta["grid"]
        tfmod = importlib.import_module(f"tf2_yolo_b200.tf_ops.{PKG[ver]}")
        ours = importlib.import_module(f"tf2_yolo_b200.{PKG[ver]}.losses")
        if ver == 2 and kw.get("anchors") is None:
            continue
        attrs = tfmod.wrap_yolo_loss((S, S), meta["B"], meta["C"], **kw).op_attrs
        t, p = torch.from_numpy(yt).cuda(), torch.from_numpy(yp).cuda()
        loss, dpred = torch.zeros((), device="cuda"), torch.zeros_like(p)
        run_op(lib, "YoloGridLoss", attrs, [t, p], [loss, dpred])
        l2, g2 = ours.wrap_yolo_loss((S, S), meta["B"], meta["C"], **kw).value_and_grad(t, p)
        assert torch.equal(loss.reshape(-1), l2.reshape(-1)) and torch.equal(dpred, g2), name
        assert abs(float(loss) - loss_ref[0]) <= 1e-5 * abs(loss_ref[0]), name
        n += 1
    assert n >= 14


def test_metrics_and_fused_ops(tf_stub):  # noqa: F811
    from tf2_yolo_b200 import synth
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    lib = build_tfstub()
    tf4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    cfg = synth.make_config("v4-608", batch=2, seed=8)
    B, Cn = 3, 80
    kws = [dict(grid_shape=(S, S), bbox_num=B, class_num=Cn, anchors=cfg["anchors"][si * B:(si + 1) * B],
                binary_weight=[1, 0.5, 2][si], loss_weight=[1, 5, 1]) for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    fns = [wrap_yolo_loss(**k) for k in kws]
    loss_ref, d_ref, _, met_ref = fused_losses(fns, yts, yps, want_metrics=True, recall_iou_threshold=0.6)
    # fused: all three scales in one launch
    fused = tf_stub.core.fused_losses([tf4.wrap_yolo_loss(**k) for k in kws])
    loss = torch.zeros(3, device="cuda")
    dps = [torch.zeros_like(p) for p in yps]
    run_op(lib, "YoloGridLossFused", dict(fused.op_attrs, N=3), yts + yps, [loss] + dps)
    assert torch.equal(loss, loss_ref) and all(torch.equal(a, b) for a, b in zip(dps, d_ref))
    # metrics, forward only (dpred comes back empty) and with gradient
    for si in range(3):
        m = tf4.wrap_recall(kws[si]["grid_shape"], B, Cn, iou_threshold=0.6)
        l1, empty, met = torch.zeros((), device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(10, dtype=torch.float64, device="cuda")
        run_op(lib, "YoloGridLossMetrics", m.op_attrs, [yts[si], yps[si]], [l1, empty, met])
        assert torch.allclose(met[:4], met_ref[si, :4], rtol=1e-12), si
        attrs = dict(tf4.wrap_yolo_loss(**kws[si]).op_attrs, recall_iou_threshold=0.6, want_grad=True)
        d1 = torch.zeros_like(yps[si])
        run_op(lib, "YoloGridLossMetrics", attrs, [yts[si], yps[si]], [l1, d1, met])
        assert torch.equal(l1, loss_ref[si]) and torch.equal(d1, d_ref[si]) and torch.equal(met, met_ref[si])


def test_compute_rejects_mismatched_tensors(tf_stub):  # noqa: F811
    lib = build_tfstub()
    tf4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    attrs = tf4.wrap_yolo_loss((4, 4), 3, 6, loss_weight=[1, 5, 1]).op_attrs
    t = torch.zeros(2 * 16 * 11, device="cuda")
    p = torch.zeros(2 * 16 * 33 + 1, device="cuda")          # not a multiple of a cell
    ins = (C.c_void_p * 2)(t.data_ptr(), p.data_ptr())
    elems = (C.c_int64 * 2)(t.numel(), p.numel())
    outs = (C.c_void_p * 2)(t.data_ptr(), p.data_ptr())
    rc = lib.tfstub_run(b"YoloGridLoss", attr_spec(attrs), 2, ins, elems, 2, outs, None, 0, None)
    assert rc == 4 and b"reshape" in lib.tfstub_last_error()


def test_from_logits_attr_reaches_the_kernel(tf_stub):  # noqa: F811
    """The op with from_logits=True == the ctypes from-logits loss on raw head outputs."""
    from tf2_yolo_b200 import synth
    from tf2_yolo_b200.grid_loss import wrap_yolo_loss_from_logits
    lib = build_tfstub()
    head = importlib.import_module("tf2_yolo_b200.tf_ops.head")
    cfg = synth.make_config("v4-608", batch=2, seed=9)
    S, B, Cn = 19, 3, 80
    anc = cfg["anchors"][:3]
    raw = torch.randn((2, S, S, B * (5 + Cn)), device="cuda") * 0.5
    yt = torch.from_numpy(cfg["y_trues"][0]).cuda()
    attrs = head.wrap_yolo_loss_from_logits(4, (S, S), B, Cn, anc, loss_weight=[1, 5, 1]).op_attrs
    loss, draw = torch.zeros((), device="cuda"), torch.zeros_like(raw)
    run_op(lib, "YoloGridLoss", attrs, [yt, raw], [loss, draw])
    l2, g2 = wrap_yolo_loss_from_logits(4, (S, S), B, Cn, anchors=anc, loss_weight=[1, 5, 1]).value_and_grad(yt, raw)
    assert torch.equal(loss.reshape(-1), l2.reshape(-1)) and torch.equal(draw, g2)
