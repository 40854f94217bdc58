"""YOLOv1 loss -- mirrors /root/reference/yolov1_5/losses/loss.py (cal_iou :9-37,
wrap_yolo_loss :40-118) on top of the fused CUDA kernel."""
from ...grid_loss import GridLoss, cal_iou_grid

EPSILON = 1e-07


def cal_iou(xywh_true, xywh_pred, grid_shape):
    return cal_iou_grid(xywh_true, xywh_pred, grid_shape)


def wrap_yolo_loss(grid_shape,
                   bbox_num,
                   class_num,
                   binary_weight=1,
                   loss_weight=[1, 1, 1, 1]):
    """Wrapped YOLOv1 loss function: returns ``yolo_loss(y_true, y_pred)``."""
    return GridLoss(1, grid_shape, bbox_num, class_num,
                    binary_weight=binary_weight, loss_weight=loss_weight)
