"""yolov3.metrics -- same names as the reference package (yolov3/metrics/yolo_metrics.py)."""
from .yolo_metrics import wrap_class_acc, wrap_mean_iou, wrap_obj_acc, wrap_recall  # noqa: F401
