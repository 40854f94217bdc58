"""CPU restatement of the four tf2_YOLO grid losses (TEST INFRASTRUCTURE ONLY).

One parametrised torch implementation (fp64 by default, autograd for
dL/dy_pred) following, term by term,

    v4   /root/reference/yolov4/losses/loss.py:10-61 (cal_iou, CIoU), :64-169
    v3   /root/reference/yolov3/losses/loss.py:9-37, :40-164
    v2   /root/reference/yolov2/losses/loss.py:9-37, :40-137
    v1   /root/reference/yolov1_5/losses/loss.py:9-37, :40-118

TensorFlow semantics that differ from stock torch are restated explicitly:
``tf.maximum``/``tf.minimum`` route the whole gradient to the first operand on
ties (torch splits it), see _max_first/_min_first.

Parity pinning: checked against the verbatim reference source executed over
the TF->torch shim (oracle/refexec.py, container only) and against
tests/golden/loss_*.npz, which that execution produced.  The reference has no
tests of its own for these functions.
"""
from dataclasses import dataclass, field
import math
from typing import Optional, Sequence

import numpy as np
import torch

EPS = 1e-07  # EPSILON in every reference loss file (loss.py:7 / :6)


class _MaxFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        pick_a = a >= b
        ctx.save_for_backward(pick_a)
        return torch.where(pick_a, a, b)

    @staticmethod
    def backward(ctx, g):
        (pick_a,) = ctx.saved_tensors
        return g * pick_a, g * (~pick_a)


class _MinFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        pick_a = a <= b
        ctx.save_for_backward(pick_a)
        return torch.where(pick_a, a, b)

    @staticmethod
    def backward(ctx, g):
        (pick_a,) = ctx.saved_tensors
        return g * pick_a, g * (~pick_a)


def _max_first(a, b):
    a, b = torch.broadcast_tensors(a, torch.as_tensor(b, dtype=a.dtype))
    return _MaxFirst.apply(a, b)


def _min_first(a, b):
    a, b = torch.broadcast_tensors(a, torch.as_tensor(b, dtype=a.dtype))
    return _MinFirst.apply(a, b)


@dataclass
class GridLossSpec:
    """Keyword surface of the four ``wrap_yolo_loss`` closures, unified."""
    version: int
    grid_shape: Sequence[int]
    bbox_num: int
    class_num: int
    anchors: Optional[Sequence[Sequence[float]]] = None
    binary_weight: float = 1.0
    loss_weight: Sequence[float] = field(default_factory=lambda: [1, 1, 1, 1])
    wh_reg_weight: float = 0.01       # v4 only (v2/v3 hard-code 0.01)
    ignore_thresh: float = 0.6
    truth_thresh: float = 1.0         # v4 only
    label_smooth: float = 0.0         # v4 only
    focal_loss_gamma: float = 2.0     # v3 (if use_focal_loss) / v4
    use_focal_loss: bool = False      # v3 only
    use_scale: bool = True            # v3 only (v2: always on)


def grid_iou(box_t, box_p, grid_hw, want_ciou=False):
    """IoU (and CIoU) between a label box and B predicted boxes of the same cell.

    ``box_*[..., :2]`` are cell-relative offsets, divided by (grid_w, grid_h);
    ``box_*[..., 2:4]`` are image-normalised sizes.  v4 loss.py:14-38 (IoU),
    :41-57 (CIoU); v1-v3 loss.py:11-35.
    """
    gh, gw = grid_hw
    scale = torch.tensor([gw, gh], dtype=box_p.dtype)
    ctr_t, size_t = box_t[..., 0:2] / scale, box_t[..., 2:4]
    ctr_p, size_p = box_p[..., 0:2] / scale, box_p[..., 2:4]
    lo_t, hi_t = ctr_t - size_t / 2.0, ctr_t + size_t / 2.0
    lo_p, hi_p = ctr_p - size_p / 2.0, ctr_p + size_p / 2.0

    ov = _max_first(_min_first(hi_p, hi_t) - _max_first(lo_p, lo_t), 0.0)
    inter = ov[..., 0] * ov[..., 1]
    union = size_p[..., 0] * size_p[..., 1] + size_t[..., 0] * size_t[..., 1] - inter
    iou = inter / (union + EPS)
    if not want_ciou:
        return iou

    hull = _max_first(hi_p, hi_t) - _min_first(lo_p, lo_t)
    diag2 = torch.pow(hull[..., 0], 2) + torch.pow(hull[..., 1], 2)
    dist2 = (torch.pow(ctr_t[..., 0] - ctr_p[..., 0], 2)
             + torch.pow(ctr_t[..., 1] - ctr_p[..., 1], 2))
    ang_t = torch.atan(size_t[..., 0] / (size_t[..., 1] + EPS))
    ang_p = torch.atan(size_p[..., 0] / (size_p[..., 1] + EPS))
    v = 4.0 / (math.pi ** 2) * torch.pow(ang_t - ang_p, 2)
    alpha = v / (1 - iou + v)           # NOT stop-gradiented (loss.py:55)
    return iou, iou - dist2 / diag2 - alpha * v


def _batch_mean_sum(x):
    """reduce_sum(reduce_mean(x, axis=0)) -- every reference term uses it."""
    return x.mean(dim=0).sum()


def loss_terms(spec: GridLossSpec, y_true, y_pred):
    """Return dict of the un-weighted terms and the weighted total (torch)."""
    gh, gw = spec.grid_shape
    B, C = spec.bbox_num, spec.class_num
    dt = y_pred.dtype
    bw = spec.binary_weight
    if isinstance(bw, np.ndarray):
        bw = torch.as_tensor(bw, dtype=dt)

    if spec.version == 1:
        return _loss_terms_v1(spec, y_true, y_pred, bw)

    ch = 5 + C
    t = y_true.reshape(-1, gh, gw, 1, ch)
    p = y_pred.reshape(-1, gh, gw, B, ch)
    anc = (torch.as_tensor(np.asarray(spec.anchors, dtype=np.float64), dtype=dt).reshape(1, 1, 1, B, 2)
           if spec.anchors is not None else 1)

    if spec.version == 4:
        iou, ciou = grid_iou(t[..., :4], p[..., :4], (gh, gw), want_ciou=True)
    else:
        iou = grid_iou(t[..., :4], p[..., :4], (gh, gw))

    best = torch.nn.functional.one_hot(iou.argmax(dim=-1), B).to(dt)
    pos = t[..., 4] * best                                   # has_obj_mask
    if spec.version == 4 and spec.truth_thresh < 1:
        pos = pos + (iou > spec.truth_thresh).to(dt) * (1 - pos)
    neg = (1 - pos) * (iou < spec.ignore_thresh).to(dt)      # no_obj_mask
    pos_e = pos.unsqueeze(-1)

    conf = p[..., 4]
    cls_t, cls_p = t[..., 5:], torch.clamp(p[..., 5:], EPS, 1 - EPS)
    log_wh_p = torch.log(p[..., 2:4] / anc)
    out = {}

    if spec.version == 4:
        out["box"] = _batch_mean_sum(pos * (1 - ciou))
        cc = torch.clamp(conf, EPS, 1 - EPS)
        if spec.label_smooth > 0:
            e_pos = torch.abs(1 - spec.label_smooth - cc)
            e_neg = torch.abs(spec.label_smooth - cc)
        else:
            e_pos, e_neg = 1 - cc, cc
        g = spec.focal_loss_gamma
        c_pos = -_batch_mean_sum(pos * e_pos ** g * torch.log(1 - e_pos))
        c_neg = -_batch_mean_sum(neg * e_neg ** g * torch.log(1 - e_neg))
        out["conf"] = c_pos + bw * c_neg
        out["cls"] = -_batch_mean_sum(
            pos_e * (cls_t * torch.log(cls_p) + (1 - cls_t) * torch.log(1 - cls_p)))
        out["reg"] = _batch_mean_sum(log_wh_p ** 2)
        lw = spec.loss_weight
        out["total"] = (lw[0] * out["box"] + lw[1] * out["conf"] + lw[2] * out["cls"]
                        + spec.wh_reg_weight * out["reg"])
        return out

    # v2 / v3
    log_wh_t = torch.log(_max_first(t[..., 2:4] / anc, EPS))
    use_scale = True if spec.version == 2 else spec.use_scale
    s = (2 - t[..., 2:3] * t[..., 3:4]) if use_scale else 1
    out["xy"] = _batch_mean_sum(pos_e * s * (t[..., 0:2] - p[..., 0:2]) ** 2)
    out["wh"] = _batch_mean_sum(pos_e * s * (log_wh_t - log_wh_p) ** 2)
    if spec.version == 3 and spec.use_focal_loss:
        cc = torch.clamp(conf, EPS, 1 - EPS)
        g = spec.focal_loss_gamma
        c_pos = -_batch_mean_sum(pos * (1 - cc) ** g * torch.log(cc))
        c_neg = -_batch_mean_sum(neg * cc ** g * torch.log(1 - cc))
    else:
        c_pos = _batch_mean_sum(pos * (1 - conf) ** 2)
        c_neg = _batch_mean_sum(neg * (0 - conf) ** 2)
    out["conf"] = c_pos + bw * c_neg
    if spec.version == 3:
        out["cls"] = -_batch_mean_sum(
            pos_e * (cls_t * torch.log(cls_p) + (1 - cls_t) * torch.log(1 - cls_p)))
    else:  # v2: positive-only cross entropy, loss.py:116-124
        out["cls"] = -_batch_mean_sum(pos_e * (cls_t * torch.log(cls_p)))
    out["reg"] = _batch_mean_sum(log_wh_p ** 2) * 0.01
    lw = spec.loss_weight
    out["total"] = (lw[0] * out["xy"] + lw[1] * out["wh"] + lw[2] * out["conf"]
                    + lw[3] * out["cls"] + out["reg"])
    return out


def _loss_terms_v1(spec, y_true, y_pred, bw):
    """v1 layout: y_pred = B x [x,y,w,h,c] then C shared class scores
    (yolov1_5/losses/loss.py:47-52, :101-102)."""
    gh, gw = spec.grid_shape
    B, C = spec.bbox_num, spec.class_num
    dt = y_pred.dtype
    t5 = y_true[..., :-C].reshape(-1, gh, gw, 1, 5)
    p5 = y_pred[..., :-C].reshape(-1, gh, gw, B, 5)
    iou = grid_iou(t5, p5, (gh, gw))
    best = torch.nn.functional.one_hot(iou.argmax(dim=-1), B).to(dt)
    obj = t5[..., 4]                                         # N,S,S,1
    resp = obj * best
    neg = 1 - resp
    conf = p5[..., 4]
    wh_t = _max_first(t5[..., 2:4], EPS)
    wh_p = _max_first(p5[..., 2:4], EPS)
    out = {}
    out["xy"] = _batch_mean_sum(resp.unsqueeze(-1) * (t5[..., 0:2] - p5[..., 0:2]) ** 2)
    out["wh"] = _batch_mean_sum(resp.unsqueeze(-1) * (torch.sqrt(wh_t) - torch.sqrt(wh_p)) ** 2)
    c_pos = _batch_mean_sum(resp * (iou - conf) ** 2)        # IoU carries gradient (:86-91)
    c_neg = _batch_mean_sum(neg * (0 - conf) ** 2)
    out["conf"] = c_pos + bw * c_neg
    cls_t = y_true[..., -C:].reshape(-1, gh, gw, C)
    cls_p = torch.clamp(y_pred[..., -C:].reshape(-1, gh, gw, C), EPS, 1 - EPS)
    out["cls"] = -_batch_mean_sum(obj * cls_t * torch.log(cls_p))
    lw = spec.loss_weight
    out["total"] = lw[0] * out["xy"] + lw[1] * out["wh"] + lw[2] * out["conf"] + lw[3] * out["cls"]
    return out


def loss_and_grad(spec: GridLossSpec, y_true, y_pred, dtype=torch.float64, threads=None):
    """(loss, dL/dy_pred, terms) as numpy, computed on CPU with autograd."""
    if threads:
        torch.set_num_threads(threads)
    yt = torch.as_tensor(np.asarray(y_true), dtype=dtype)
    yp = torch.as_tensor(np.asarray(y_pred), dtype=dtype).clone().requires_grad_(True)
    terms = loss_terms(spec, yt, yp)
    total = terms["total"]
    total.reshape(-1)[0].backward() if total.numel() == 1 else total.sum().backward()
    grad = yp.grad.detach().numpy()
    return (float(total.detach().reshape(-1)[0]), grad,
            {k: float(v.detach().reshape(-1)[0]) for k, v in terms.items()})


def activate_head(raw, anchors, bbox_num, class_num):
    """The v3/v4 head transform (yolov4/models/__init__.py:42-60, Anchor layer backbone.py:59-60):
    sigmoid on xy / objectness / class scores, anchor * exp on wh; torch, differentiable."""
    shp = raw.shape
    z = raw.reshape(-1, bbox_num, 5 + class_num)
    anc = torch.as_tensor(np.asarray(anchors, dtype=np.float64), dtype=raw.dtype).reshape(1, bbox_num, 2)
    out = torch.cat([torch.sigmoid(z[..., 0:2]), torch.exp(z[..., 2:4]) * anc, torch.sigmoid(z[..., 4:])], dim=-1)
    return out.reshape(shp)


def loss_and_grad_from_logits(spec: GridLossSpec, y_true, raw, dtype=torch.float64):
    """(loss, dL/d raw): the oracle loss composed with the head transform."""
    yt = torch.as_tensor(np.asarray(y_true), dtype=dtype)
    z = torch.as_tensor(np.asarray(raw), dtype=dtype).clone().requires_grad_(True)
    act = activate_head(z, spec.anchors, spec.bbox_num, spec.class_num)
    total = loss_terms(spec, yt, act)["total"]
    total.reshape(-1)[0].backward()
    return float(total.detach().reshape(-1)[0]), z.grad.detach().numpy()
