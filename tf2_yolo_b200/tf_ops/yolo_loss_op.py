"""TensorFlow binding of the fused grid loss (``tf.load_op_library`` + registered gradient).

Only importable where TensorFlow and the compiled ``libyolo_b200_tf.so`` exist (see the build
line in yolo_loss_op.cc).  In the B200 build container TensorFlow is absent, so this module is
source-only there and the same C entry points are exercised through ctypes
(tf2_yolo_b200/engine.py); importing it without TensorFlow raises ImportError - no fallback.

Usage in the reference (yolov4/__init__.py:523-535 builds the loss list)::

    from tf2_yolo_b200.tf_ops.yolo_loss_op import wrap_yolo_loss   # instead of yolov4.losses
    model.compile(optimizer, loss=[wrap_yolo_loss(grid_shape=..., bbox_num=3, class_num=80, ...) ...])
"""
import os

import numpy as np
import tensorflow as tf  # noqa: F401  (ImportError here is the intended failure mode)
from tensorflow.python.framework import ops

_lib = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libyolo_b200_tf.so"))


@ops.RegisterGradient("YoloGridLoss")
def _yolo_grid_loss_grad(op, grad_loss, _grad_dpred):
    # dL/dy_pred was produced by the forward kernel in the same pass; labels get no gradient.
    return None, grad_loss * op.outputs[1]


def wrap_yolo_loss(version=4, *, grid_shape, bbox_num, class_num, anchors=None, binary_weight=1,
                   loss_weight=(1, 1, 1), wh_reg_weight=0.01, ignore_thresh=.6, truth_thresh=1,
                   label_smooth=0, focal_loss_gamma=2, use_focal_loss=False, use_scale=True,
                   global_batch=0):
    anchors_flat = [] if anchors is None else [float(v) for v in np.asarray(anchors).reshape(-1)]
    bw = float(np.asarray(binary_weight).reshape(-1)[0])
    out_shape = np.shape(binary_weight) if isinstance(binary_weight, np.ndarray) else ()

    def yolo_loss(y_true, y_pred):
        loss, _ = _lib.yolo_grid_loss(
            y_true=tf.cast(y_true, tf.float32), y_pred=y_pred, version=version,
            grid_h=int(grid_shape[0]), grid_w=int(grid_shape[1]), bbox_num=bbox_num, class_num=class_num,
            anchors=anchors_flat, binary_weight=bw, loss_weight=[float(w) for w in loss_weight],
            wh_reg_weight=wh_reg_weight, ignore_thresh=ignore_thresh, truth_thresh=truth_thresh,
            label_smooth=label_smooth, focal_loss_gamma=focal_loss_gamma,
            use_focal_loss=use_focal_loss, use_scale=use_scale, global_batch=global_batch)
        return tf.reshape(loss, out_shape)

    return yolo_loss
