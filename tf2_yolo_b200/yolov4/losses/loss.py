"""YOLOv4 loss -- mirrors /root/reference/yolov4/losses/loss.py (cal_iou :10-61,
wrap_yolo_loss :64-169) on top of the fused CUDA kernel."""
from ...grid_loss import GridLoss, cal_iou_grid

EPSILON = 1e-07


def cal_iou(xywh_true, xywh_pred, grid_shape, return_ciou=False):
    """IoU (and CIoU) of label boxes vs predicted boxes; shape (N, S, S, B)[, x2]."""
    return cal_iou_grid(xywh_true, xywh_pred, grid_shape, return_ciou=return_ciou)


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
    return GridLoss(4, grid_shape, bbox_num, class_num,
                    anchors=anchors, binary_weight=binary_weight, loss_weight=loss_weight,
                    wh_reg_weight=wh_reg_weight, ignore_thresh=ignore_thresh,
                    truth_thresh=truth_thresh, label_smooth=label_smooth,
                    focal_loss_gamma=focal_loss_gamma)
