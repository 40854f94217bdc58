"""CPU: the TensorFlow boundary (tf2_yolo_b200/tf_ops) checked WITHOUT TensorFlow.

* the Python bindings run against a recording stand-in for ``tensorflow`` and are replayed with the
  exact keyword sets of the reference's four call sites (Yolo.loss / Yolo.metrics);
* the C++ op source is compiled against tests/tf_stub (a functional stand-in for the slice of the
  TF op API it uses) and its REGISTER_OP attribute lists are cross-checked against what the Python
  side passes; kernel construction (attribute validation) runs for every call site.
The GPU half (Compute against libyolo_b200.so) is tests/test_gpu_tf_op.py.
"""
import importlib
import inspect
import re
import sys
import types

import numpy as np
import pytest

from conftest import attr_spec, build_tfstub


class _Recorder:
    """Fake op library: every op call records its keyword arguments."""

    def __init__(self):
        self.calls = []

    def _op(self, name, n_out):
        def call(**kw):
            self.calls.append((name, kw))
            if name == "yolo_grid_loss_fused":
                return ["loss"] + ["dpred"] * len(kw["y_pred"])
            return tuple([np.float32(1.5), "dpred", np.arange(10.0)][:n_out])
        return call

    def __getattr__(self, name):
        return self._op(name, {"yolo_grid_loss": 2, "yolo_grid_loss_metrics": 3}.get(name, 1))


@pytest.fixture()
def tf_stub(monkeypatch):
    rec = _Recorder()
    grads = {}
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.cast = lambda x, dt: x
    tf.reshape = lambda x, shape: (x, tuple(shape))
    tf.load_op_library = lambda path: rec
    tf.loaded_from = None
    py = types.ModuleType("tensorflow.python")
    fw = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")

    def RegisterGradient(name):
        def deco(f):
            grads[name] = f
            return f
        return deco
    ops.RegisterGradient = RegisterGradient
    fw.ops = ops
    py.framework = fw
    tf.python = py
    for name, mod in (("tensorflow", tf), ("tensorflow.python", py), ("tensorflow.python.framework", fw),
                      ("tensorflow.python.framework.ops", ops)):
        monkeypatch.setitem(sys.modules, name, mod)
    for m in [k for k in sys.modules if k.startswith("tf2_yolo_b200.tf_ops")]:
        monkeypatch.delitem(sys.modules, m)
    core = importlib.import_module("tf2_yolo_b200.tf_ops.yolo_loss_op")
    yield types.SimpleNamespace(rec=rec, grads=grads, core=core)
    for m in [k for k in sys.modules if k.startswith("tf2_yolo_b200.tf_ops")]:
        sys.modules.pop(m, None)


ANCHORS9 = [[0.1 * i + 0.05, 0.1 * i + 0.07] for i in range(9)]

# keyword sets exactly as the reference's Yolo.loss passes them
CALL_SITES = {
    # yolov4/__init__.py:523-535
    "yolov4": dict(grid_shape=(38, 38), bbox_num=3, class_num=80, anchors=ANCHORS9[3:6], binary_weight=1,
                   loss_weight=[1, 5, 1], wh_reg_weight=0.01, ignore_thresh=0.6, truth_thresh=1.0, label_smooth=0.0,
                   focal_loss_gamma=2),
    # yolov3/__init__.py:425-436
    "yolov3": dict(grid_shape=(26, 26), bbox_num=3, class_num=80, anchors=ANCHORS9[3:6], binary_weight=1,
                   loss_weight=[1, 1, 5, 1], ignore_thresh=.6, use_focal_loss=False, focal_loss_gamma=2,
                   use_scale=True),
    # yolov2/__init__.py:311-318
    "yolov2": dict(grid_shape=(13, 13), bbox_num=5, class_num=20, anchors=ANCHORS9[:5], binary_weight=1,
                   loss_weight=[1, 1, 5, 1], ignore_thresh=0.6),
    # yolov1_5/__init__.py:291-297
    "yolov1_5": dict(grid_shape=(7, 7), bbox_num=2, class_num=20, binary_weight=0.5, loss_weight=[5, 5, 1, 1]),
}
VERSION = {"yolov4": 4, "yolov3": 3, "yolov2": 2, "yolov1_5": 1}


def test_signatures_equal_the_reference(tf_stub):
    from test_host_logic import REF_SIGNATURES
    for (pkg, fn), names in REF_SIGNATURES.items():
        mod = importlib.import_module(f"tf2_yolo_b200.tf_ops.{pkg}")
        assert list(inspect.signature(getattr(mod, fn)).parameters) == names, pkg
        ours = importlib.import_module(f"tf2_yolo_b200.{pkg}.losses")
        # same defaults as the ctypes mirror (which test_host_logic pins to the reference)
        a = {k: v.default for k, v in inspect.signature(getattr(mod, fn)).parameters.items()}
        b = {k: v.default for k, v in inspect.signature(getattr(ours, fn)).parameters.items()}
        assert a == b, pkg
    m1 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov1_5")
    assert list(inspect.signature(m1.wrap_class_acc).parameters) == ["grid_shape", "class_num"]
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    assert list(inspect.signature(m4.wrap_recall).parameters) == ["grid_shape", "bbox_num", "class_num", "iou_threshold"]


@pytest.mark.parametrize("pkg", list(CALL_SITES))
def test_reference_call_sites_reach_the_op_with_the_right_attrs(tf_stub, pkg):
    mod = importlib.import_module(f"tf2_yolo_b200.tf_ops.{pkg}")
    kw = CALL_SITES[pkg]
    loss_fn = mod.wrap_yolo_loss(**kw)                    # keywords only, as the reference calls it
    out = loss_fn("YT", "YP")
    name, got = tf_stub.rec.calls[-1]
    assert name == "yolo_grid_loss"
    assert got["version"] == VERSION[pkg]                  # round-1 bug: stayed 4 for every package
    assert (got["grid_h"], got["grid_w"]) == tuple(kw["grid_shape"])
    assert got["bbox_num"] == kw["bbox_num"] and got["class_num"] == kw["class_num"]
    assert got["loss_weight"] == [float(w) for w in kw["loss_weight"]]
    assert got["binary_weight"] == float(kw["binary_weight"])
    assert got["anchors"] == ([] if pkg == "yolov1_5" else [float(v) for a in kw["anchors"] for v in a])
    assert got["y_true"] == "YT" and got["y_pred"] == "YP"
    assert got["from_logits"] is False and got["global_batch"] == 0
    for k in ("wh_reg_weight", "ignore_thresh", "truth_thresh", "label_smooth", "focal_loss_gamma"):
        if k in kw:
            assert got[k] == float(kw[k]), k
    for k in ("use_focal_loss", "use_scale"):
        if k in kw:
            assert got[k] is bool(kw[k])
    assert out == (np.float32(1.5), ())                    # scalar shape for a Python-float binary_weight


def test_defaults_per_version_and_array_binary_weight(tf_stub):
    # defaults: (1,1,1,1) for v1-v3, (1,1,1) for v4 (ADVICE: a 3-entry default dropped the class term)
    for pkg, n in (("yolov1_5", 4), ("yolov3", 4), ("yolov4", 3)):
        mod = importlib.import_module(f"tf2_yolo_b200.tf_ops.{pkg}")
        f = mod.wrap_yolo_loss((8, 8), 2, 3)
        assert f.op_attrs["loss_weight"] == [1.0] * n and f.op_attrs["version"] == VERSION[pkg]
    m2 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov2")
    with pytest.raises(TypeError):
        m2.wrap_yolo_loss((13, 13), 5, 20)                 # anchors has no default in yolov2
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    f = m4.wrap_yolo_loss((19, 19), 3, 80, binary_weight=np.array([0.25]))   # get_class_weight's 1-element array
    assert f("a", "b")[1] == (1,) and f.op_attrs["binary_weight"] == 0.25
    with pytest.raises(IndexError):
        importlib.import_module("tf2_yolo_b200.tf_ops.yolov3").wrap_yolo_loss((8, 8), 3, 4, loss_weight=[1, 1, 1])
    with pytest.raises(ValueError):
        tf_stub.core.make_loss(7, (8, 8), 3, 4)


def test_metrics_wrappers(tf_stub):
    # yolov4/__init__.py:556-590: positional (grid_shape, abox_num, class_num[, iou_threshold=])
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    fns = [m4.wrap_obj_acc((19, 19), 3, 80), m4.wrap_mean_iou((19, 19), 3, 80), m4.wrap_class_acc((19, 19), 3, 80),
           m4.wrap_recall((19, 19), 3, 80, iou_threshold=0.6)]
    assert [f.__name__ for f in fns] == ["obj_acc", "mean_iou", "class_acc", "recall"]
    for i, f in enumerate(fns):
        assert f("t", "p") == float(i)                     # element i of the metrics output
        name, got = tf_stub.rec.calls[-1]
        assert name == "yolo_grid_loss_metrics" and got["version"] == 4 and got["want_grad"] is False
    assert tf_stub.rec.calls[-1][1]["recall_iou_threshold"] == 0.6
    m1 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov1_5")
    yp = types.SimpleNamespace(shape=(4, 7, 7, 2 * 5 + 20))
    m1.wrap_class_acc((7, 7), 20)("t", yp)
    assert tf_stub.rec.calls[-1][1]["bbox_num"] == 2 and tf_stub.rec.calls[-1][1]["version"] == 1


def test_fused_scales_and_registered_gradients(tf_stub):
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    fns = [m4.wrap_yolo_loss(grid_shape=(19 * 2 ** i, 19 * 2 ** i), bbox_num=3, class_num=80,
                             anchors=ANCHORS9[3 * i:3 * i + 3], binary_weight=[1, 2, 3][i], loss_weight=[1, 5, 1])
           for i in range(3)]
    fused = tf_stub.core.fused_losses(fns)
    assert fused(["a", "b", "c"], ["p", "q", "r"]) == "loss"
    name, got = tf_stub.rec.calls[-1]
    assert name == "yolo_grid_loss_fused"
    assert got["grid_h"] == [19, 38, 76] and got["binary_weight"] == [1.0, 2.0, 3.0]
    assert got["anchors"] == [float(v) for a in ANCHORS9 for v in a]
    assert got["y_true"] == ["a", "b", "c"] and got["y_pred"] == ["p", "q", "r"]
    assert set(tf_stub.grads) == {"YoloGridLoss", "YoloGridLossMetrics", "YoloGridLossFused"}
    op = types.SimpleNamespace(outputs=[None, np.array([1.0, 2.0])])
    none, g = tf_stub.grads["YoloGridLoss"](op, 3.0, None)
    assert none is None and np.array_equal(g, [3.0, 6.0])   # (None, upstream * dpred)
    op3 = types.SimpleNamespace(outputs=[None, np.array([1.0]), np.array([2.0]), np.array([3.0])])
    out = tf_stub.grads["YoloGridLossFused"](op3, np.array([10.0, 20.0, 30.0]), None, None, None)
    assert out[:3] == [None, None, None] and [float(x[0]) for x in out[3:]] == [10.0, 40.0, 90.0]


# ---- the C++ side ---------------------------------------------------------------------------------
def registered_ops():
    lib = build_tfstub()
    ops = {}
    for line in lib.tfstub_describe_ops().decode().strip().split("\n"):
        name, ins, outs, attrs, device = line.split("|")
        ops[name] = dict(inputs=[x for x in ins.split(";") if x], outputs=[x for x in outs.split(";") if x],
                         attrs={a.split(":")[0].strip(): a.split(":", 1)[1].strip() for a in attrs.split(";") if a},
                         device=device)
    return lib, ops


def _tf_type(v):
    if isinstance(v, bool):
        return "bool"
    if isinstance(v, int):
        return "int"
    if isinstance(v, float):
        return "float"
    return None


def test_cc_compiles_and_its_attr_lists_match_the_python_side(tf_stub):
    lib, ops = registered_ops()
    assert set(ops) == {"YoloGridLoss", "YoloGridLossMetrics", "YoloGridLossFused"}
    assert all(o["device"] == "GPU" for o in ops.values())          # no CPU kernel registered
    # the source text agrees with the registry (the test reads what was compiled, not a copy)
    src = open(build_tfstub.__globals__["ROOT"] + "/tf2_yolo_b200/tf_ops/yolo_loss_op.cc").read()
    assert len(re.findall(r"REGISTER_OP\(", src)) == 3 and len(re.findall(r"REGISTER_KERNEL_BUILDER\(", src)) == 3
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    loss_attrs = m4.wrap_yolo_loss(**CALL_SITES["yolov4"]).op_attrs
    metric_attrs = m4.wrap_recall((19, 19), 3, 80).op_attrs
    fns = [m4.wrap_yolo_loss(grid_shape=(19 << i, 19 << i), bbox_num=3, class_num=80, anchors=ANCHORS9[3 * i:3 * i + 3],
                             loss_weight=[1, 5, 1]) for i in range(3)]
    fused_attrs = dict(tf_stub.core.fused_losses(fns).op_attrs, N=3)   # N is inferred by TF from the input lists
    for op, passed in (("YoloGridLoss", loss_attrs), ("YoloGridLossMetrics", metric_attrs),
                       ("YoloGridLossFused", fused_attrs)):
        declared = ops[op]["attrs"]
        assert set(passed) == set(declared), (op, set(passed) ^ set(declared))
        for k, v in passed.items():
            t = declared[k].split("=")[0].split(">")[0].strip()
            if t.startswith("list("):
                assert isinstance(v, list), (op, k)
                inner = t[5:-1]
                assert all(_tf_type(x) == inner or (inner == "float" and isinstance(x, float)) for x in v), (op, k, v)
            else:
                assert _tf_type(v) == t, (op, k, v, t)


@pytest.mark.parametrize("pkg", list(CALL_SITES))
def test_kernel_construction_accepts_every_reference_call_site(tf_stub, pkg):
    lib, _ = registered_ops()
    mod = importlib.import_module(f"tf2_yolo_b200.tf_ops.{pkg}")
    attrs = mod.wrap_yolo_loss(**CALL_SITES[pkg]).op_attrs
    assert lib.tfstub_construct(b"YoloGridLoss", attr_spec(attrs)) == 0, lib.tfstub_last_error()
    kw = CALL_SITES[pkg]
    rec = mod.wrap_recall(kw["grid_shape"], kw["bbox_num"], kw["class_num"], iou_threshold=0.6).op_attrs
    assert lib.tfstub_construct(b"YoloGridLossMetrics", attr_spec(rec)) == 0, lib.tfstub_last_error()


def test_kernel_construction_rejects_what_the_reference_cannot_mean(tf_stub):
    lib, _ = registered_ops()
    m4 = importlib.import_module("tf2_yolo_b200.tf_ops.yolov4")
    good = m4.wrap_yolo_loss(**CALL_SITES["yolov4"]).op_attrs
    for bad, msg in ((dict(good, version=7), b"version"),
                     (dict(good, loss_weight=[1.0, 1.0, 1.0, 1.0]), b"loss_weight"),       # v4 takes 3
                     (dict(good, version=3), b"loss_weight"),                              # v3 takes 4
                     (dict(good, anchors=[0.1, 0.2]), b"anchors"),
                     (dict(good, version=2, loss_weight=[1.0] * 4, anchors=[]), b"anchors"),
                     (dict(good, bbox_num=40), b"bbox_num"),
                     (dict(good, version=2, loss_weight=[1.0] * 4, from_logits=True), b"from_logits")):
        assert lib.tfstub_construct(b"YoloGridLoss", attr_spec(bad)) == 1
        assert msg in lib.tfstub_last_error(), (bad, lib.tfstub_last_error())
    missing = dict(good)
    del missing["class_num"]
    assert lib.tfstub_construct(b"YoloGridLoss", attr_spec(missing)) == 1


def test_raw_logit_head_keeps_the_reference_layer_names(tf_stub, monkeypatch):
    """yolo_head_logits builds the reference head's convolutions under the reference's names
    (yolov4/models/__init__.py:42-66) with no activation and no Anchor layer, in the per-box order
    [xy, wh, conf, prob] the from_logits kernel reads; its loss reaches the op with from_logits."""
    made = []

    def conv(filters, size, activation=None, name=None):
        made.append((name, filters, size, activation))
        return lambda t: f"{name}({t})"
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Concatenate = lambda name=None: (lambda parts: (name, list(parts)))
    models = types.ModuleType("tensorflow.keras.models")
    models.Model = lambda inp, outs: types.SimpleNamespace(input=inp, output=outs)
    keras = types.ModuleType("tensorflow.keras")
    for name, mod in (("tensorflow.keras", keras), ("tensorflow.keras.layers", layers), ("tensorflow.keras.models", models)):
        monkeypatch.setitem(sys.modules, name, mod)
    head = importlib.import_module("tf2_yolo_b200.tf_ops.head")
    body = types.SimpleNamespace(input="img", output=["f19", "f38", "f76"])
    model = head.yolo_head_logits(body, class_num=80, anchors=ANCHORS9, conv_layer=conv)
    assert model.input == "img" and [o[0] for o in model.output] == ["out1_concat", "out2_concat", "out3_concat"]
    assert len(made) == 3 * 3 * 4 and all(m[3] is None for m in made)          # 36 convolutions, none activated
    assert [m[0] for m in made[:4]] == ["out1_box1_xy_conv", "out1_box1_wh_conv", "out1_box1_conf_conv",
                                         "out1_box1_prob_conv"]
    assert [m[1] for m in made[:4]] == [2, 2, 1, 80]
    assert model.output[2][1][4].startswith("out3_box2_xy_conv(f76")               # per-box order inside a scale
    with pytest.raises(ValueError):
        head.yolo_head_logits(body, anchors=ANCHORS9[:8], conv_layer=conv)
    f = head.wrap_yolo_loss_from_logits(4, (19, 19), 3, 80, ANCHORS9[:3], loss_weight=[1, 5, 1])
    assert f.op_attrs["from_logits"] is True and f.op_attrs["version"] == 4
    lib, _ = registered_ops()
    assert lib.tfstub_construct(b"YoloGridLoss", attr_spec(f.op_attrs)) == 0, lib.tfstub_last_error()
    with pytest.raises(ValueError):
        head.wrap_yolo_loss_from_logits(2, (13, 13), 5, 20, ANCHORS9[:5])
