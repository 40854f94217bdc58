"""Mirror of /root/reference/utils/kmeans.py (iou :9-24, iou_dist :27-33,
euclidean_dist :36-40, kmeans :43-102).

The Lloyd loop keeps the reference's control flow and its use of the global
``numpy.random`` stream (initial centres; re-draws of empty clusters, ascending
cluster index) on the host.  Each iteration's distance matrix / argmin /
per-cluster mean (kmeans.py:79-90) is ONE streaming CUDA kernel
(yb_kmeans_assign) over the device-resident boxes; only k*(d+1) numbers come
back per iteration.  With ``process_group`` the boxes are sharded over ranks and
those k*(d+1) partial sums are all-reduced (NCCL) before the centres update.
"""
import numpy as np
import torch
from numpy.random import rand

from .. import dist as dist_util
from .. import engine
from .._native import YB_DIST_EUCLID, YB_DIST_IOU, YoloB200Error

_HOST_LIMIT = 1 << 16  # the direct-call helpers below are for centre-sized arrays only


def _small(*arrays):
    for a in arrays:
        if np.size(a) > _HOST_LIMIT:
            raise YoloB200Error("distance helpers are host-side for centre-sized inputs only; "
                                "use kmeans() for data-sized work (no CPU fallback)")


def iou(center_boxes, data_boxes):
    """Area ratio min/max (the reference's ``iou``; not a box overlap)."""
    _small(center_boxes, data_boxes)
    ca = center_boxes[..., 0] * center_boxes[..., 1]
    da = data_boxes[..., 0] * data_boxes[..., 1]
    return np.minimum(ca, da) / np.maximum(ca, da)


def iou_dist(center_boxes, data_boxes):
    return 1 - iou(center_boxes, data_boxes)


def euclidean_dist(center_boxes, data_boxes):
    _small(center_boxes, data_boxes)
    return np.sqrt(np.sum(np.square(center_boxes - data_boxes), axis=-1))


_KIND = {iou_dist: YB_DIST_IOU, euclidean_dist: YB_DIST_EUCLID}


def kmeans(data, n_cluster, dist_func, stop_dist, max_iternum=10000, verbose=True,
           process_group=None, return_assignments=False):
    """k-means with the reference's semantics; ``data`` is (num_samples, num_dims)
    as an ndarray or a float64 CUDA tensor (this rank's shard when sharded)."""
    if dist_func not in _KIND:
        raise YoloB200Error("dist_func must be this module's iou_dist or euclidean_dist "
                            "(arbitrary Python distance functions cannot run on the GPU)")
    kind = _KIND[dist_func]
    if torch.is_tensor(data):
        dev_data = data.to(torch.float64).contiguous()
        if not dev_data.is_cuda:
            dev_data = dev_data.cuda()
    else:
        if not torch.cuda.is_available():
            raise YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
        dev_data = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64))).cuda()
    n_dim = dev_data.shape[-1]
    dev_data = dev_data.reshape(-1, n_dim)
    mm = engine.minmax(dev_data)
    if process_group is not None:
        mm = dist_util.allreduce_minmax(mm, process_group)
    data_min, data_max = (float(v) for v in mm.cpu().numpy())
    data_min, data_max = np.float64(data_min), np.float64(data_max)

    center = rand(n_cluster * n_dim).reshape((n_cluster, 1, n_dim)) * data_max
    center = center * (data_max - data_min) + data_min

    epoch = 1
    assign = None
    while True:
        dev_center = torch.from_numpy(np.ascontiguousarray(center.reshape(n_cluster, n_dim))).to(dev_data.device)
        assign, sums, counts = engine.kmeans_assign(dev_data, dev_center, kind, want_assign=return_assignments)
        if process_group is not None:
            sums, counts = dist_util.allreduce_kmeans(sums, counts, process_group)
        sums = sums.cpu().numpy()
        counts = counts.cpu().numpy()
        new_center = np.copy(center)
        for n in range(n_cluster):
            if counts[n] > 0:
                cluster = sums[n] / counts[n]
            else:
                cluster = rand(n_dim) * (data_max - data_min) + data_min
            new_center[n, 0] = cluster

        loss = np.mean(dist_func(center, new_center))
        center = new_center
        if verbose:
            print(f"epoch {epoch:2d}: loss = {loss:.4f}")
        epoch += 1
        if loss < stop_dist or epoch > max_iternum:
            break

    center = center.reshape((n_cluster, n_dim))
    center = center.astype("float32")
    if return_assignments:
        return center, assign
    return center
