"""YOLOv2 loss -- mirrors /root/reference/yolov2/losses/loss.py (cal_iou :9-37,
wrap_yolo_loss :40-137) on top of the fused CUDA kernel."""
from ...grid_loss import GridLoss, cal_iou_grid

EPSILON = 1e-07


def cal_iou(xywh_true, xywh_pred, grid_shape):
    return cal_iou_grid(xywh_true, xywh_pred, grid_shape)


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   anchors,
                   binary_weight=1,
                   loss_weight=[1, 1, 1, 1],
                   ignore_thresh=.6):
    """Wrapped YOLOv2 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return GridLoss(2, grid_shape, bbox_num, class_num,
                    anchors=anchors, binary_weight=binary_weight, loss_weight=loss_weight,
                    ignore_thresh=ignore_thresh)
