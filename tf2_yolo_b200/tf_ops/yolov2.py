"""TensorFlow-side drop-in for ``yolov2.losses`` / ``yolov2.metrics`` of the reference: the
signatures of yolov2/losses/loss.py:40-46 and yolov2/metrics/yolo_metrics.py, bound to version 2 of the
YoloGridLoss custom op (tf_ops/yolo_loss_op.cc).  The call site it serves is ``Yolo.loss`` /
``Yolo.metrics`` (yolov2/__init__.py:311-318), which passes keywords only::

    # yolov2/__init__.py
    -from .losses import wrap_yolo_loss
    +from tf2_yolo_b200.tf_ops.yolov2 import wrap_yolo_loss
"""
from .yolo_loss_op import make_loss, make_metric

VERSION = 2


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   anchors,
                   binary_weight=1,
                   loss_weight=[1, 1, 1, 1],
                   ignore_thresh=.6):
    """Wrapped YOLOv2 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return make_loss(2, grid_shape, bbox_num, class_num,
                     anchors=anchors, binary_weight=binary_weight, loss_weight=loss_weight,
                     ignore_thresh=ignore_thresh)


def wrap_obj_acc(grid_shape, bbox_num, class_num):
    """Wrapped objectness accuracy."""
    return make_metric(2, "obj_acc", grid_shape, bbox_num, class_num)


def wrap_mean_iou(grid_shape, bbox_num, class_num):
    """Wrapped mean IoU."""
    return make_metric(2, "mean_iou", grid_shape, bbox_num, class_num)


def wrap_class_acc(grid_shape, bbox_num, class_num):
    """Wrapped class accuracy."""
    return make_metric(2, "class_acc", grid_shape, bbox_num, class_num)


def wrap_recall(grid_shape, bbox_num, class_num, iou_threshold=0.5):
    """Wrapped bounding box recall."""
    return make_metric(2, "recall", grid_shape, bbox_num, class_num, iou_threshold)
