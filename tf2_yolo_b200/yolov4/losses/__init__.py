"""yolov4.losses -- same names as the reference package."""
from .loss import cal_iou, wrap_yolo_loss  # noqa: F401
