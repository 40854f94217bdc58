"""TensorFlow binding of the fused grid loss: ``tf.load_op_library`` + registered gradients.

Only importable where TensorFlow and the compiled ``libyolo_b200_tf.so`` exist (build line in
yolo_loss_op.cc).  In the B200 build container TensorFlow is absent, so this module is exercised
there with a recording stand-in for ``tensorflow`` (tests/test_tf_ops.py) and the C++ op source is
compiled and run against a stub of the TF op API (tests/tf_stub/); the same C entry points are
reached through ctypes by the rest of the package.  Importing it without TensorFlow raises
ImportError - there is no fallback.

The reference-facing wrappers live next to this file, one module per reference package with the
reference's own signatures::

    tf2_yolo_b200.tf_ops.yolov4   wrap_yolo_loss, wrap_obj_acc, wrap_mean_iou, wrap_class_acc, wrap_recall
    tf2_yolo_b200.tf_ops.yolov3   (same names)        tf2_yolo_b200.tf_ops.yolov2, .yolov1_5

``make_loss(version, ...)`` below is what they all call; ``version`` is a required positional
argument, never a default.
"""
import os

import numpy as np
import tensorflow as tf  # noqa: F401  (ImportError here is the intended failure mode)
from tensorflow.python.framework import ops

_lib = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libyolo_b200_tf.so"))

METRIC_INDEX = {"obj_acc": 0, "mean_iou": 1, "class_acc": 2, "recall": 3}
# loss_weight defaults of the four reference signatures (yolov1_5/losses/loss.py:44,
# yolov2/...:45, yolov3/...:45, yolov4/...:69)
DEFAULT_LOSS_WEIGHT = {1: (1, 1, 1, 1), 2: (1, 1, 1, 1), 3: (1, 1, 1, 1), 4: (1, 1, 1)}


@ops.RegisterGradient("YoloGridLoss")
def _yolo_grid_loss_grad(op, grad_loss, _grad_dpred):
    # dL/dy_pred was produced by the forward kernel in the same pass; labels get no gradient.
    return None, grad_loss * op.outputs[1]


@ops.RegisterGradient("YoloGridLossMetrics")
def _yolo_grid_loss_metrics_grad(op, grad_loss, _grad_dpred, _grad_metrics):
    return None, grad_loss * op.outputs[1]


@ops.RegisterGradient("YoloGridLossFused")
def _yolo_grid_loss_fused_grad(op, grad_loss, *_grad_dpreds):
    n = (len(op.outputs) - 1)
    return [None] * n + [grad_loss[i] * op.outputs[1 + i] for i in range(n)]


def _attrs(version, grid_shape, bbox_num, class_num, anchors, binary_weight, loss_weight, wh_reg_weight,
           ignore_thresh, truth_thresh, label_smooth, focal_loss_gamma, use_focal_loss, use_scale, from_logits,
           global_batch):
    """Keyword arguments of a reference ``wrap_yolo_loss`` call -> attribute dict of the op."""
    if version not in (1, 2, 3, 4):
        raise ValueError(f"Invalid version: {version}")
    if loss_weight is None:
        loss_weight = DEFAULT_LOSS_WEIGHT[version]
    loss_weight = [float(w) for w in loss_weight]
    need = len(DEFAULT_LOSS_WEIGHT[version])
    if len(loss_weight) < need:
        raise IndexError("list index out of range")       # what indexing loss_weight[need-1] raises in the reference
    if version == 2 and anchors is None:
        raise TypeError("wrap_yolo_loss() missing 1 required positional argument: 'anchors'")
    anchors_flat = [] if (anchors is None or version == 1) else [float(v) for v in np.asarray(anchors).reshape(-1)]
    return dict(
        version=int(version), grid_h=int(grid_shape[0]), grid_w=int(grid_shape[1]), bbox_num=int(bbox_num),
        class_num=int(class_num), anchors=anchors_flat,
        binary_weight=float(np.asarray(binary_weight, dtype=np.float64).reshape(-1)[0]),
        loss_weight=loss_weight[:need], wh_reg_weight=float(wh_reg_weight), ignore_thresh=float(ignore_thresh),
        truth_thresh=float(truth_thresh), label_smooth=float(label_smooth),
        focal_loss_gamma=float(focal_loss_gamma), use_focal_loss=bool(use_focal_loss), use_scale=bool(use_scale),
        from_logits=bool(from_logits), global_batch=int(global_batch))


def make_loss(version, grid_shape, bbox_num, class_num, anchors=None, binary_weight=1, loss_weight=None,
              wh_reg_weight=0.01, ignore_thresh=.6, truth_thresh=1, label_smooth=0, focal_loss_gamma=2,
              use_focal_loss=False, use_scale=True, from_logits=False, global_batch=0):
    """``yolo_loss(y_true, y_pred)`` backed by the YoloGridLoss op.  ``global_batch`` = 0 divides
    by the batch the op sees (what the reference's reduce_mean(axis=0) does; under
    ``tf.distribute`` Keras scales the per-replica losses itself); a positive value is the divisor
    for hand-written multi-replica loops that sum the per-replica losses."""
    attrs = _attrs(version, grid_shape, bbox_num, class_num, anchors, binary_weight, loss_weight, wh_reg_weight,
                   ignore_thresh, truth_thresh, label_smooth, focal_loss_gamma, use_focal_loss, use_scale,
                   from_logits, global_batch)
    # a 1-element ndarray binary_weight (utils/tools.py:616-620) makes the reference return shape (1,)
    out_shape = np.shape(binary_weight) if isinstance(binary_weight, np.ndarray) else ()

    def yolo_loss(y_true, y_pred):
        loss, _ = _lib.yolo_grid_loss(y_true=tf.cast(y_true, tf.float32), y_pred=y_pred, **attrs)
        return tf.reshape(loss, out_shape)

    yolo_loss.op_attrs = attrs
    return yolo_loss


def make_metric(version, kind, grid_shape, bbox_num, class_num, iou_threshold=0.5):
    """One of the four in-training metrics (yolov*/metrics/yolo_metrics.py) from the
    YoloGridLossMetrics op: the metrics ride along with a forward pass of the loss kernel."""
    attrs = _attrs(version, grid_shape, bbox_num, class_num, None if version != 2 else [1.0] * (2 * bbox_num), 1,
                   None, 0.01, .6, 1, 0, 2, False, True, False, 0)
    attrs["recall_iou_threshold"] = float(iou_threshold)
    attrs["want_grad"] = False
    index = METRIC_INDEX[kind]

    def metric(y_true, y_pred):
        _, _, m = _lib.yolo_grid_loss_metrics(y_true=tf.cast(y_true, tf.float32), y_pred=y_pred, **attrs)
        return tf.cast(m[index], tf.float32)

    metric.__name__ = kind
    metric.op_attrs = attrs
    return metric


def fused_losses(loss_fns):
    """One launch for the FPN outputs of a train step: ``fused(y_trues, y_preds) -> loss [N]``
    from the closures ``Yolo.loss()`` returned (yolov4/__init__.py:518-535, yolov3/__init__.py:408-424)."""
    a = [f.op_attrs for f in loss_fns]
    same = ("version", "bbox_num", "class_num", "loss_weight", "wh_reg_weight", "ignore_thresh", "truth_thresh",
            "label_smooth", "focal_loss_gamma", "use_focal_loss", "use_scale", "from_logits", "global_batch")
    for k in same:
        if any(x[k] != a[0][k] for x in a):
            raise ValueError(f"fused scales must share {k}")
    attrs = {k: a[0][k] for k in same}
    attrs.update(grid_h=[x["grid_h"] for x in a], grid_w=[x["grid_w"] for x in a],
                 binary_weight=[x["binary_weight"] for x in a],
                 anchors=[v for x in a for v in x["anchors"]])

    def fused(y_trues, y_preds):
        out = _lib.yolo_grid_loss_fused(y_true=[tf.cast(t, tf.float32) for t in y_trues], y_pred=list(y_preds),
                                        **attrs)
        return out[0]

    fused.op_attrs = attrs
    return fused
