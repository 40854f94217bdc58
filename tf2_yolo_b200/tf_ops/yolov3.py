"""TensorFlow-side drop-in for ``yolov3.losses`` / ``yolov3.metrics`` of the reference: the
signatures of yolov3/losses/loss.py:40-49 and yolov3/metrics/yolo_metrics.py, bound to version 3 of the
YoloGridLoss custom op (tf_ops/yolo_loss_op.cc).  The call site it serves is ``Yolo.loss`` /
``Yolo.metrics`` (yolov3/__init__.py:425-436), which passes keywords only::

    # yolov3/__init__.py
    -from .losses import wrap_yolo_loss
    +from tf2_yolo_b200.tf_ops.yolov3 import wrap_yolo_loss
"""
from .yolo_loss_op import make_loss, make_metric

VERSION = 3


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   anchors=None,
                   binary_weight=1,
                   loss_weight=[1, 1, 1, 1],
                   ignore_thresh=.6,
                   use_focal_loss=False,
                   focal_loss_gamma=2,
                   use_scale=True):
    """Wrapped YOLOv3 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return make_loss(3, grid_shape, bbox_num, class_num,
                     anchors=anchors, binary_weight=binary_weight, loss_weight=loss_weight,
                     ignore_thresh=ignore_thresh, use_focal_loss=use_focal_loss,
                     focal_loss_gamma=focal_loss_gamma, use_scale=use_scale)


def wrap_obj_acc(grid_shape, bbox_num, class_num):
    """Wrapped objectness accuracy."""
    return make_metric(3, "obj_acc", grid_shape, bbox_num, class_num)


def wrap_mean_iou(grid_shape, bbox_num, class_num):
    """Wrapped mean IoU."""
    return make_metric(3, "mean_iou", grid_shape, bbox_num, class_num)


def wrap_class_acc(grid_shape, bbox_num, class_num):
    """Wrapped class accuracy."""
    return make_metric(3, "class_acc", grid_shape, bbox_num, class_num)


def wrap_recall(grid_shape, bbox_num, class_num, iou_threshold=0.5):
    """Wrapped bounding box recall."""
    return make_metric(3, "recall", grid_shape, bbox_num, class_num, iou_threshold)
