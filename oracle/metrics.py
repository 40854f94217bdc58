"""CPU restatement of the in-training metrics (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/yolov{2,3,4}/metrics/yolo_metrics.py: obj_acc :9-28, mean_iou :31-55,
class_acc :58-82, recall :85-115 and the v1 layout variant yolov1_5/metrics/yolo_metrics.py.
Keras averages a per-sample metric tensor, so obj_acc here is the mean over all cells.

Parity pinning: checked against the verbatim reference files executed over the TF shim
(oracle/refexec.reference_metrics) and tests/golden/metrics.npz.
"""
import numpy as np
import torch

from .losses import grid_iou

EPS = 1e-07


def grid_metrics(version, y_true, y_pred, grid_shape, bbox_num, class_num, iou_threshold=0.5,
                 dtype=torch.float64):
    """dict(obj_acc, mean_iou, class_acc, recall) plus the raw sums."""
    gh, gw = grid_shape
    B, C = bbox_num, class_num
    yt = torch.as_tensor(np.asarray(y_true), dtype=dtype)
    yp = torch.as_tensor(np.asarray(y_pred), dtype=dtype)
    if version == 1:
        t5 = yt[..., :-C].reshape(-1, gh, gw, 1, 5)
        p5 = yp[..., :-C].reshape(-1, gh, gw, B, 5)
        cls_t = yt[..., -C:].reshape(-1, gh, gw, 1, C)
        cls_p = yp[..., -C:].reshape(-1, gh, gw, 1, C)
    else:
        t = yt.reshape(-1, gh, gw, 1, 5 + C)
        p = yp.reshape(-1, gh, gw, B, 5 + C)
        t5, p5, cls_t, cls_p = t[..., :5], p[..., :5], t[..., 5:], p[..., 5:]
    obj = t5[..., 4]                                                   # N,S,S,1
    conf_max = p5[..., 4].max(dim=-1, keepdim=True).values
    hits = (obj == (conf_max > 0.5).to(dtype)).to(dtype).sum()
    iou = grid_iou(t5[..., :4], p5[..., :4], (gh, gw))                 # N,S,S,B
    sum_iou = (iou.max(dim=-1, keepdim=True).values * obj).sum()
    n_obj = obj.sum()
    equal = (cls_t.argmax(dim=-1) == cls_p.argmax(dim=-1)).to(dtype) * obj   # N,S,S,B (v1: N,S,S,1)
    sum_eq = equal.sum()
    tp = ((iou * equal).max(dim=-1, keepdim=True).values >= iou_threshold).to(dtype).sum()
    cells = float(obj.numel())
    denom_cls = n_obj * (1 if version == 1 else B)
    out = {"obj_acc": float(hits) / cells, "mean_iou": float(sum_iou / (n_obj + EPS)),
           "class_acc": float(sum_eq / (denom_cls + EPS)), "recall": float(tp / (n_obj + EPS)),
           "raw": [float(hits), float(sum_iou), float(n_obj), float(sum_eq), float(tp), cells]}
    return out
