"""TensorFlow-side drop-in for ``yolov1_5.losses`` / ``yolov1_5.metrics`` of the reference: the
signatures of yolov1_5/losses/loss.py:40-44 and yolov1_5/metrics/yolo_metrics.py, bound to version 1 of the
YoloGridLoss custom op (tf_ops/yolo_loss_op.cc).  The call site it serves is ``Yolo.loss`` /
``Yolo.metrics`` (yolov1_5/__init__.py:291-297), which passes keywords only::

    # yolov1_5/__init__.py
    -from .losses import wrap_yolo_loss
    +from tf2_yolo_b200.tf_ops.yolov1_5 import wrap_yolo_loss
"""
from .yolo_loss_op import make_loss, make_metric

VERSION = 1


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   binary_weight=1,
                   loss_weight=[1, 1, 1, 1]):
    """Wrapped YOLOv1 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return make_loss(1, grid_shape, bbox_num, class_num,
                     binary_weight=binary_weight, loss_weight=loss_weight)


def wrap_obj_acc(grid_shape, bbox_num, class_num):
    """Wrapped objectness accuracy."""
    return make_metric(1, "obj_acc", grid_shape, bbox_num, class_num)


def wrap_mean_iou(grid_shape, bbox_num, class_num):
    """Wrapped mean IoU."""
    return make_metric(1, "mean_iou", grid_shape, bbox_num, class_num)


def wrap_class_acc(grid_shape, class_num):
    """Wrapped class accuracy (yolov1_5/metrics/yolo_metrics.py:52: per cell, no bbox_num)."""
    def class_acc(y_true, y_pred):
        bbox_num = (int(y_pred.shape[-1]) - class_num) // 5
        return make_metric(1, "class_acc", grid_shape, bbox_num, class_num)(y_true, y_pred)
    return class_acc


def wrap_recall(grid_shape, bbox_num, class_num, iou_threshold=0.5):
    """Wrapped bounding box recall."""
    return make_metric(1, "recall", grid_shape, bbox_num, class_num, iou_threshold)
