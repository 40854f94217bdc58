"""Mirror of the hot-path functions of /root/reference/utils/tools.py:
``decode`` (:370-438), ``cal_iou`` (:630-684), ``nms`` (:687-733) with the
reference's NumPy signatures and return types, executed on the GPU through the
C ABI (yb_decode / yb_pairwise_iou / yb_nms).  ``decode_batch`` / ``nms_batch``
are the device-resident, whole-batch forms the per-image functions are built on.
"""
import numpy as np
import torch

from .. import engine
from .._native import YoloB200Error

EPSILON = 1e-07


def _device():
    if not torch.cuda.is_available():
        raise YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_grid(a, dev, dtype=None):
    if torch.is_tensor(a):
        t = a
    else:
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64 if a.dtype.itemsize > 4 else np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev, non_blocking=True).contiguous()


decode_batch = engine.decode_batch_exact
nms_batch = engine.nms_batch


def encode_labels(boxes, box_offsets, img_size, grid_shape, class_num, n_levels=1, dtype=np.float64):
    """Label grids of a batch from its box lists: what ``YoloDataSequence.__getitem__`` builds in
    its ``_encode_to_array`` closure (utils/tools.py:179-209), plus - for ``n_levels`` > 1 - the
    ``down2xlabel`` pyramid of ``_Yolov4DataSequence`` (yolov4/__init__.py:47-53), coarse grid
    first.  ``boxes``: (n_boxes, 5) [x1, y1, x2, y2, class index] in pixels of the resized image,
    ``box_offsets``: (n_img+1,) first box of every image, ``img_size`` = (height, width),
    ``grid_shape`` = the finest grid.  Returns a list of float64 (or ``dtype``) ndarrays.
    Raises IndexError / ValueError where the reference would (non-finite corners, unknown class,
    cell index below -grid)."""
    dev = _device()
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 5))
    off = np.ascontiguousarray(np.asarray(box_offsets, dtype=np.int64).reshape(-1))
    if off.size < 1 or off[0] != 0 or off[-1] != b.shape[0] or np.any(np.diff(off) < 0):
        raise ValueError("box_offsets must rise from 0 to len(boxes)")
    most = int(np.diff(off).max()) if off.size > 1 else 0
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    outs, n_bad = engine.encode_labels(torch.from_numpy(b).to(dev), torch.from_numpy(off).to(dev), img_size,
                                       grid_shape, class_num, n_levels, tdt, max_boxes_per_img=most)
    res = [o.cpu().numpy() for o in outs]
    if int(n_bad.item()):
        if not np.all(np.isfinite(b[:, :4])):
            raise ValueError("cannot convert float NaN to integer")       # int(nan // w) in the reference
        raise IndexError("index out of bounds for the label grid (box class or cell index)")
    return res


def down2xlabel(label_data):
    """Downsample label by 2x (utils/tools.py:342-367): ndarray in, float64 ndarray out."""
    dev = _device()
    a = np.asarray(label_data)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return engine.down2x_labels(t).cpu().numpy()


def get_class_weight(label_data, method="alpha"):
    """Weight of every category (utils/tools.py:592-627).  The per-class sums over the label
    tensor run on the GPU; the arithmetic on the C sums follows the reference on the host."""
    dev = _device()
    a = label_data if torch.is_tensor(label_data) else np.asarray(label_data)
    n_cls = a.shape[-1]
    if torch.is_tensor(a):
        t = a.to(dev)
        t = t if t.dtype in (torch.float32, torch.float64) else t.double()
    else:
        a = a if a.dtype in (np.float32, np.float64) else a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sums = engine.column_sums(t.reshape(-1, n_cls).contiguous()).cpu().numpy()
    class_weight = []
    if method != "alpha":
        total = 1
        for i in a.shape[:-1]:
            total *= i
        if method == "effective":
            beta = (total - 1)/total
    for i in range(n_cls):
        samples_per_class = sums[i]
        if method == "effective":
            effective_num = 1 - np.power(beta, samples_per_class)
            class_weight.append((1 - beta)/effective_num)
        elif method == "binary":
            class_weight.append(samples_per_class/(total - samples_per_class))
        else:
            class_weight.append(1/samples_per_class)
    class_weight = np.array(class_weight)
    if method == "log":
        class_weight = np.log(total*class_weight)

    if method != "binary":
        class_weight = class_weight/np.sum(class_weight)*len(class_weight)

    return class_weight


def decode(*label_datas, class_num=1, threshold=0.5, version=1):
    """Decode the prediction (or label) grids of ONE image.

    Args and return value as the reference: ndarrays (grid_h, grid_w, info) in,
    ``(K, 7)`` float64 rows [x, y, w, h, c, class index, class probability] out;
    shape ``(0,)`` when nothing passes the threshold.
    """
    if version not in (1, 2, 3, 4):
        raise ValueError(f"Invalid version: {version}")
    if len(label_datas) == 0:
        return np.array([], dtype="float")
    dev = _device()
    f64 = any((torch.is_tensor(a) and a.dtype == torch.float64) or
              (not torch.is_tensor(a) and np.asarray(a).dtype == np.float64) for a in label_datas)
    dt = torch.float64 if f64 else torch.float32
    chunks = []
    # scales share one launch when they fit the ABI's scale table
    for i in range(0, len(label_datas), 4):
        grids = [_as_device_grid(a, dev, dt).unsqueeze(0) for a in label_datas[i:i + 4]]
        rows, _ = engine.decode_batch_exact(grids, class_num, threshold, version)
        chunks.append(rows)
    rows = torch.cat(chunks, dim=0) if len(chunks) > 1 else chunks[0]
    out = rows.cpu().numpy()
    if out.shape[0] == 0:
        return np.array([], dtype="float")
    return out


def cal_iou(xywh_true, xywh_pred, mode=1):
    """IoU (mode 1) / DIoU (mode 2) with NumPy broadcasting semantics."""
    if mode not in (1, 2):
        return None  # the reference falls off the end of the function
    a = np.asarray(xywh_true, dtype=np.float64)
    b = np.asarray(xywh_pred, dtype=np.float64)
    dev = _device()
    lead = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    if (a.ndim == 3 and b.ndim == 3 and a.shape[1] == 1 and b.shape[0] == 1):
        ta = torch.from_numpy(np.ascontiguousarray(a[:, 0, :])).to(dev)
        tb = torch.from_numpy(np.ascontiguousarray(b[0, :, :])).to(dev)
        return engine.pairwise_iou(ta, tb, mode).cpu().numpy()
    ab = np.ascontiguousarray(np.broadcast_to(a, lead + a.shape[-1:])).reshape(-1, a.shape[-1])
    bb = np.ascontiguousarray(np.broadcast_to(b, lead + b.shape[-1:])).reshape(-1, b.shape[-1])
    out = engine.elementwise_iou(torch.from_numpy(ab).to(dev), torch.from_numpy(bb).to(dev), mode)
    return out.cpu().numpy().reshape(lead)


def nms(xywhcp, class_num=1, nms_threshold=0.45, iou_mode=1):
    """Per-class greedy NMS (iou_mode 1) / DIoU-NMS (iou_mode 2) of one image's rows."""
    xywhcp = np.asarray(xywhcp, dtype=np.float64)
    if xywhcp.ndim != 2:
        raise IndexError("too many indices for array: nms needs (K, 7) rows from decode()")
    dev = _device()
    rows = torch.from_numpy(np.ascontiguousarray(xywhcp)).to(dev)
    offsets = torch.tensor([0, rows.shape[0]], dtype=torch.int64, device=dev)
    r = engine.nms_batch(rows, offsets, class_num, nms_threshold, iou_mode)
    n = int(r["out_offsets"][-1].item())
    return r["out_rows"][:n].cpu().numpy()


def soft_nms(xywhcp, class_num=1,
        nms_threshold=0.45, conf_threshold=0.5, sigma=0.5):
    """Gaussian Soft-NMS of one image's rows (utils/tools.py:736-786)."""
    xywhcp = np.asarray(xywhcp, dtype=np.float64)
    if xywhcp.ndim != 2:
        raise IndexError("too many indices for array: soft_nms needs (K, 7) rows from decode()")
    dev = _device()
    rows = torch.from_numpy(np.ascontiguousarray(xywhcp)).to(dev)
    offsets = torch.tensor([0, rows.shape[0]], dtype=torch.int64, device=dev)
    r = engine.nms_batch(rows, offsets, class_num, nms_threshold, 1, soft=(conf_threshold, sigma))
    n = int(r["out_offsets"][-1].item())
    return r["out_rows"][:n].cpu().numpy()
