"""Execute the UNMODIFIED reference source (container only; test infrastructure).

The reference (/root/reference) is pure Python on TensorFlow/NumPy.  Its NumPy
half imports here once seven absent third-party imports are stubbed; its TF
half (the four ``losses/loss.py`` files) uses 16 TensorFlow symbols, which a
small shim maps onto torch so the files run verbatim with autograd providing
dL/dy_pred.

This module exists to (a) generate tests/golden/ and (b) cross-check the
restatements in oracle/{losses,tools,kmeans,measurement}.py.  It cannot travel
to the GPU box (no /root/reference there); everything that runs on the box
uses the restatements plus the committed fixtures.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("YB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils"))


# --------------------------------------------------------------------------
# TensorFlow -> torch shim with TF gradient tie rules
# --------------------------------------------------------------------------
class _TFMax(torch.autograd.Function):
    """tf.maximum: on ties the whole gradient goes to the FIRST argument
    (TF's MaximumGrad uses x >= y)."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x >= y)
        return torch.maximum(x, y)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return g * m, g * (~m)


class _TFMin(torch.autograd.Function):
    """tf.minimum: MinimumGrad uses x <= y."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x <= y)
        return torch.minimum(x, y)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return g * m, g * (~m)


def _bcast2(fn):
    def wrapped(x, y):
        ref = x if torch.is_tensor(x) else y
        x = torch.as_tensor(x, dtype=ref.dtype)
        y = torch.as_tensor(y, dtype=ref.dtype)
        xb, yb = torch.broadcast_tensors(x, y)
        return fn(xb, yb)  # autograd handles un-broadcast of expanded views
    return wrapped


def make_tf_shim():
    tf = types.ModuleType("tensorflow")
    tf.reshape = lambda x, shape: torch.reshape(torch.as_tensor(x), tuple(int(s) for s in shape))
    tf.maximum = _bcast2(_TFMax.apply)
    tf.minimum = _bcast2(_TFMin.apply)
    tf.pow = lambda x, p: torch.pow(x, p)
    tf.atan = torch.atan
    tf.square = lambda x: x * x
    tf.sqrt = torch.sqrt
    tf.argmax = lambda x, axis=-1: torch.argmax(x, dim=axis)
    tf.one_hot = lambda i, depth, dtype=None: torch.nn.functional.one_hot(i, depth).to(dtype)
    tf.cast = lambda x, dtype: x.to(dtype)
    tf.expand_dims = lambda x, axis: torch.unsqueeze(x, axis)
    tf.reduce_sum = lambda x, axis=None: x.sum() if axis is None else x.sum(dim=axis)
    tf.reduce_mean = lambda x, axis=None: x.mean() if axis is None else x.mean(dim=axis)
    tf.clip_by_value = lambda x, lo, hi: torch.clamp(x, lo, hi)
    tf.math = types.SimpleNamespace(log=torch.log, abs=torch.abs)
    return tf


class _TorchNP:
    """Stand-in for the module-global ``np`` of the loss files: ``np.array`` must
    return something a grad-requiring torch tensor can be divided by."""

    def __init__(self, dtype):
        self._dtype = dtype

    def array(self, x):
        return torch.tensor(list(x), dtype=self._dtype)


_STUBS = {
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "matplotlib.patches": {"Rectangle": object, "Circle": object, "BoxStyle": object},
    "bs4": {"BeautifulSoup": object},
    "imgaug": {},
    "imgaug.augmentables": {},
    "imgaug.augmentables.bbs": {"BoundingBox": object, "BoundingBoxesOnImage": object},
    "tensorflow.keras": {},
    "tensorflow.keras.utils": {"Sequence": object},
    "cv2": {},
}


def _install_stubs():
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            if name.split(".")[0] in ("cv2",):
                importlib.import_module(name)
                continue
        except Exception:
            pass
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = make_tf_shim()
    try:
        import PIL  # noqa: F401
    except Exception:
        pil = types.ModuleType("PIL")
        pil.Image = types.ModuleType("PIL.Image")
        sys.modules["PIL"] = pil
        sys.modules["PIL.Image"] = pil.Image


_np_half = None


def load_numpy_half():
    """Return (tools, kmeans, measurement) modules of the reference, unmodified."""
    global _np_half
    if _np_half is None:
        if not available():
            raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
        _install_stubs()
        mods = {}
        pkg = types.ModuleType("yb_ref_utils")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "utils")]
        sys.modules["yb_ref_utils"] = pkg
        for name in ("tools", "kmeans", "measurement"):
            spec = importlib.util.spec_from_file_location(
                "yb_ref_utils." + name, os.path.join(REFERENCE_ROOT, "utils", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["yb_ref_utils." + name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
        _np_half = (mods["tools"], mods["kmeans"], mods["measurement"])
    return _np_half


def load_loss_module(version, dtype=torch.float64):
    """Load yolov{1_5,2,3,4}/losses/loss.py verbatim over the TF shim."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _install_stubs()
    pkg = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}[version]
    path = os.path.join(REFERENCE_ROOT, pkg, "losses", "loss.py")
    spec = importlib.util.spec_from_file_location(f"yb_ref_{pkg}_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.tf = make_tf_shim()
    mod.np = _TorchNP(dtype)
    return mod


def reference_loss(version, y_true, y_pred, dtype=torch.float64, **wrap_kwargs):
    """loss, dL/dy_pred of the verbatim reference closure, on CPU."""
    mod = load_loss_module(version, dtype)
    kw = dict(wrap_kwargs)
    if kw.get("anchors") is not None:
        kw["anchors"] = torch.tensor(np.asarray(kw["anchors"], dtype=np.float64), dtype=dtype)
    if isinstance(kw.get("binary_weight"), np.ndarray):
        # TF multiplies ndarray * Tensor by converting the ndarray; torch needs help.
        kw["binary_weight"] = torch.tensor(kw["binary_weight"], dtype=dtype)
    fn = mod.wrap_yolo_loss(**kw)
    yt = torch.as_tensor(np.asarray(y_true), dtype=dtype)
    yp = torch.as_tensor(np.asarray(y_pred), dtype=dtype).clone().requires_grad_(True)
    loss = fn(yt, yp)
    loss = loss.reshape(()) if loss.numel() == 1 else loss
    loss.backward()
    return loss.detach().numpy().copy(), yp.grad.detach().numpy().copy()


def load_metrics_module(version, dtype=torch.float64):
    """Load yolov{1_5,2,3,4}/metrics/yolo_metrics.py verbatim over the TF shim (its
    ``binary_accuracy`` import and ``from yolovX.losses import cal_iou`` are satisfied by shim
    modules; cal_iou is the reference's own, from the verbatim loss file)."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _install_stubs()
    pkg = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}[version]
    loss_mod = load_loss_module(version, dtype)
    tf = make_tf_shim()
    tf.reduce_max = lambda x, axis=None, keepdims=False: (x.max() if axis is None
                                                          else x.max(dim=axis, keepdim=keepdims).values)
    km = types.ModuleType("tensorflow.keras.metrics")
    km.binary_accuracy = lambda y_true, y_pred, threshold=0.5: (
        (y_true == (y_pred > threshold).to(y_pred.dtype)).to(y_pred.dtype).mean(dim=-1))
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow.keras.metrics", pkg, pkg + ".losses")}
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras.metrics"] = km
    fake_pkg = types.ModuleType(pkg)
    fake_pkg.__path__ = []
    fake_losses = types.ModuleType(pkg + ".losses")
    fake_losses.cal_iou = loss_mod.cal_iou
    sys.modules[pkg] = fake_pkg
    sys.modules[pkg + ".losses"] = fake_losses
    try:
        path = os.path.join(REFERENCE_ROOT, pkg, "metrics", "yolo_metrics.py")
        spec = importlib.util.spec_from_file_location(f"yb_ref_{pkg}_metrics", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.tf = tf
    return mod


def reference_metrics(version, y_true, y_pred, grid_shape, bbox_num, class_num, iou_threshold=0.5,
                      dtype=torch.float64):
    """[obj_acc, mean_iou, class_acc, recall] of the verbatim reference closures (Keras reduces a
    per-sample metric tensor by its mean, which is what obj_acc needs)."""
    mod = load_metrics_module(version, dtype)
    yt = torch.as_tensor(np.asarray(y_true), dtype=dtype)
    yp = torch.as_tensor(np.asarray(y_pred), dtype=dtype)
    fns = [mod.wrap_obj_acc(grid_shape, bbox_num, class_num),
           mod.wrap_mean_iou(grid_shape, bbox_num, class_num),
           mod.wrap_class_acc(grid_shape, class_num) if version == 1 else mod.wrap_class_acc(grid_shape, bbox_num, class_num),
           mod.wrap_recall(grid_shape, bbox_num, class_num, iou_threshold=iou_threshold)]
    return np.array([float(torch.as_tensor(f(yt, yp), dtype=dtype).mean()) for f in fns])


class _RefBox:
    """Stand-in for imgaug's BoundingBox: the reader only stores the four corners and
    ``_encode_to_array`` (utils/tools.py:190-194) only reads them back."""

    def __init__(self, x1, y1, x2, y2):
        self.x1, self.y1, self.x2, self.y2 = x1, y1, x2, y2


class _RefBoxes:
    def __init__(self, bounding_boxes, shape=None):
        self.bounding_boxes = bounding_boxes
        self.shape = shape


def reference_encode_labels(boxes, box_offsets, img_size, grid_shape, class_num):
    """Label grid produced by the UNMODIFIED ``YoloDataSequence.__getitem__`` (utils/tools.py:176-339)
    reading labelme files written to a temporary directory: blank PNGs of exactly ``img_size`` (so the
    reader's zoom ratio is 1.0 and the corners reach ``_encode_to_array`` unchanged) and one JSON per
    image whose rectangles are ``boxes`` ([x1, y1, x2, y2, class index], pixel units)."""
    import json
    import tempfile

    from PIL import Image

    tools, _, _ = load_numpy_half()
    tools.BoundingBox, tools.BoundingBoxesOnImage = _RefBox, _RefBoxes
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 5)
    off = np.asarray(box_offsets, dtype=np.int64)
    n_img = len(off) - 1
    names = [f"c{k}" for k in range(class_num)]
    with tempfile.TemporaryDirectory() as tmp:
        img_dir, lab_dir = os.path.join(tmp, "img"), os.path.join(tmp, "lab")
        os.makedirs(img_dir)
        os.makedirs(lab_dir)
        blank = Image.fromarray(np.zeros((int(img_size[0]), int(img_size[1]), 3), dtype=np.uint8))
        for i in range(n_img):
            blank.save(os.path.join(img_dir, f"im{i:05d}.png"))
            shapes = [dict(label=names[int(b[4])], shape_type="rectangle",
                           points=[[float(b[0]), float(b[1])], [float(b[2]), float(b[3])]])
                      for b in boxes[off[i]:off[i + 1]]]
            with open(os.path.join(lab_dir, f"im{i:05d}.json"), "w", encoding="big5") as f:
                json.dump(dict(shapes=shapes), f)
        seq = tools.YoloDataSequence(img_path=img_dir, label_path=lab_dir, reader="PIL", batch_size=max(n_img, 1),
                                     label_format="labelme", size=(int(img_size[0]), int(img_size[1])),
                                     rescale=None, grid_shape=tuple(grid_shape), class_names=names,
                                     shuffle=False, thread_num=1)
        _, label_data = seq[0]
    return label_data
