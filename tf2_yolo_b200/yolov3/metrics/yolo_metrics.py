"""In-training metrics of yolov3 -- mirrors /root/reference/yolov3/metrics/yolo_metrics.py; computed by
the fused CUDA loss kernel's metric accumulators (see tf2_yolo_b200/grid_metrics.py)."""
from ...grid_metrics import make_module_functions

EPSILON = 1e-07

wrap_obj_acc, wrap_mean_iou, wrap_class_acc, wrap_recall = make_module_functions(3)
