"""TensorFlow-side drop-in for ``yolov4.losses`` / ``yolov4.metrics`` of the reference: the
signatures of yolov4/losses/loss.py:64-74 and yolov4/metrics/yolo_metrics.py, bound to version 4 of the
YoloGridLoss custom op (tf_ops/yolo_loss_op.cc).  The call site it serves is ``Yolo.loss`` /
``Yolo.metrics`` (yolov4/__init__.py:523-535), which passes keywords only::

    # yolov4/__init__.py
    -from .losses import wrap_yolo_loss
    +from tf2_yolo_b200.tf_ops.yolov4 import wrap_yolo_loss
"""
from .yolo_loss_op import make_loss, make_metric

VERSION = 4


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   anchors=None,
                   binary_weight=1,
                   loss_weight=[1, 1, 1],
                   wh_reg_weight=0.01,
                   ignore_thresh=.6,
                   truth_thresh=1,
                   label_smooth=0,
                   focal_loss_gamma=2):
    """Wrapped YOLOv4 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return make_loss(4, grid_shape, bbox_num, class_num,
                     anchors=anchors, binary_weight=binary_weight, loss_weight=loss_weight,
                     wh_reg_weight=wh_reg_weight, ignore_thresh=ignore_thresh, truth_thresh=truth_thresh,
                     label_smooth=label_smooth, focal_loss_gamma=focal_loss_gamma)


def wrap_obj_acc(grid_shape, bbox_num, class_num):
    """Wrapped objectness accuracy."""
    return make_metric(4, "obj_acc", grid_shape, bbox_num, class_num)


def wrap_mean_iou(grid_shape, bbox_num, class_num):
    """Wrapped mean IoU."""
    return make_metric(4, "mean_iou", grid_shape, bbox_num, class_num)


def wrap_class_acc(grid_shape, bbox_num, class_num):
    """Wrapped class accuracy."""
    return make_metric(4, "class_acc", grid_shape, bbox_num, class_num)


def wrap_recall(grid_shape, bbox_num, class_num, iou_threshold=0.5):
    """Wrapped bounding box recall."""
    return make_metric(4, "recall", grid_shape, bbox_num, class_num, iou_threshold)
