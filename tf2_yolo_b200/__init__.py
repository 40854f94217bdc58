"""tf2_yolo_b200 -- the anchor-grid hot path of samson6460/tf2_YOLO as sm_100a CUDA kernels
behind the reference's own Python API.

    tf2_yolo_b200.yolov{1_5,2,3,4}.losses.wrap_yolo_loss / cal_iou
    tf2_yolo_b200.yolov{1_5,2,3,4}.metrics.wrap_obj_acc / wrap_mean_iou / wrap_class_acc / wrap_recall
    tf2_yolo_b200.utils.tools.decode / nms / cal_iou          (+ decode_batch / nms_batch)
    tf2_yolo_b200.utils.kmeans.kmeans / iou_dist / euclidean_dist
    tf2_yolo_b200.utils.measurement.PRfunc / create_score_mat

Everything computes through libyolo_b200.so (C ABI, include/yolo_b200.h); there is no CPU or
PyTorch fallback.  ``install()`` rebinds the reference's own modules to these functions.
"""
import importlib
import sys

__all__ = ["install", "HOT_PATH"]

# (reference module, attribute) -> (module in this package, attribute)
HOT_PATH = {
    ("utils.tools", "decode"): ("tf2_yolo_b200.utils.tools", "decode"),
    ("utils.tools", "nms"): ("tf2_yolo_b200.utils.tools", "nms"),
    ("utils.tools", "cal_iou"): ("tf2_yolo_b200.utils.tools", "cal_iou"),
    ("utils.tools", "soft_nms"): ("tf2_yolo_b200.utils.tools", "soft_nms"),
    ("utils.tools", "down2xlabel"): ("tf2_yolo_b200.utils.tools", "down2xlabel"),
    ("utils.tools", "get_class_weight"): ("tf2_yolo_b200.utils.tools", "get_class_weight"),
    ("utils.measurement", "soft_nms"): ("tf2_yolo_b200.utils.tools", "soft_nms"),
    ("utils.kmeans", "kmeans"): ("tf2_yolo_b200.utils.kmeans", "kmeans"),
    ("utils.kmeans", "iou_dist"): ("tf2_yolo_b200.utils.kmeans", "iou_dist"),
    ("utils.kmeans", "euclidean_dist"): ("tf2_yolo_b200.utils.kmeans", "euclidean_dist"),
    ("utils.kmeans", "iou"): ("tf2_yolo_b200.utils.kmeans", "iou"),
    ("utils.measurement", "decode"): ("tf2_yolo_b200.utils.tools", "decode"),
    ("utils.measurement", "nms"): ("tf2_yolo_b200.utils.tools", "nms"),
    ("utils.measurement", "cal_iou"): ("tf2_yolo_b200.utils.tools", "cal_iou"),
    ("utils.measurement", "create_score_mat"): ("tf2_yolo_b200.utils.measurement", "create_score_mat"),
    ("utils.measurement", "PRfunc"): ("tf2_yolo_b200.utils.measurement", "PRfunc"),
    ("utils.measurement", "PR_func"): ("tf2_yolo_b200.utils.measurement", "PR_func"),
    ("yolov4.losses", "wrap_yolo_loss"): ("tf2_yolo_b200.yolov4.losses", "wrap_yolo_loss"),
    ("yolov4.losses", "cal_iou"): ("tf2_yolo_b200.yolov4.losses", "cal_iou"),
    ("yolov3.losses", "wrap_yolo_loss"): ("tf2_yolo_b200.yolov3.losses", "wrap_yolo_loss"),
    ("yolov3.losses", "cal_iou"): ("tf2_yolo_b200.yolov3.losses", "cal_iou"),
    ("yolov2.losses", "wrap_yolo_loss"): ("tf2_yolo_b200.yolov2.losses", "wrap_yolo_loss"),
    ("yolov2.losses", "cal_iou"): ("tf2_yolo_b200.yolov2.losses", "cal_iou"),
    ("yolov1_5.losses", "wrap_yolo_loss"): ("tf2_yolo_b200.yolov1_5.losses", "wrap_yolo_loss"),
    ("yolov1_5.losses", "cal_iou"): ("tf2_yolo_b200.yolov1_5.losses", "cal_iou"),
}
# in-training metrics: yolov*/metrics/yolo_metrics.py, imported by the facades
# as `from .metrics import wrap_obj_acc, ...` (yolov4/__init__.py:29-30)
for _pkg in ("yolov1_5", "yolov2", "yolov3", "yolov4"):
    for _fn in ("wrap_obj_acc", "wrap_mean_iou", "wrap_class_acc", "wrap_recall"):
        HOT_PATH[(f"{_pkg}.metrics", _fn)] = (f"tf2_yolo_b200.{_pkg}.metrics", _fn)


def install(modules=None):
    """Rebind the hot-path names of the reference's modules (already imported, found in
    ``sys.modules`` or passed as {name: module}) to the CUDA-backed functions.  Packages that
    did ``from .losses import wrap_yolo_loss`` (yolov4/__init__.py:24) keep their own binding,
    so their ``wrap_yolo_loss`` global is rebound as well.  Returns the list of rebound names."""
    modules = dict(modules) if modules is not None else sys.modules
    done = []
    for (ref_mod, attr), (our_mod, our_attr) in HOT_PATH.items():
        target = getattr(importlib.import_module(our_mod), our_attr)
        for name in (ref_mod, ref_mod.split(".")[0] if ref_mod.endswith((".losses", ".metrics")) else None):
            m = modules.get(name) if name else None
            if m is not None and hasattr(m, attr) and name != our_mod and not name.startswith("tf2_yolo_b200"):
                setattr(m, attr, target)
                done.append(f"{name}.{attr}")
    return done
