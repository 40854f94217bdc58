"""Mirror of /root/reference/utils/kmeans.py (iou :9-24, iou_dist :27-33,
euclidean_dist :36-40, kmeans :43-102).

The Lloyd loop runs on the GPU without the host in it: every iteration is ONE launch over the
device-resident boxes (yb_kmeans_lloyd_step: distance / argmin / per-cluster sums, and - by the
last CTA of the same launch - the update of kmeans.py:84-97 with the loss in NumPy's summation
order and the stop test).  A finished loop freezes itself on the device, so iterations are queued
in batches and the host looks at the loop state once per batch; the centres do not depend on the
batch size.  The host owns what the reference's host owns: the global ``numpy.random`` stream -
initial centres (kmeans.py:73) and the re-draw of empty clusters in ascending cluster index
(kmeans.py:89), for which the device hands the iteration back.  With ``process_group`` the boxes
are sharded over ranks: each iteration all-reduces k*(d+1) doubles on the same stream (no host
sync) and every rank applies the same update; random draws come from rank 0 and are broadcast.
"""
import numpy as np
import torch
from numpy.random import rand

from .. import dist as dist_util
from .. import engine
from .._native import YB_DIST_EUCLID, YB_DIST_IOU, YoloB200Error

_HOST_LIMIT = 1 << 16  # below this the distance helpers are centre-sized bookkeeping (host NumPy)


def _device():
    if not torch.cuda.is_available():
        raise YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _on_device(what, a, b):
    """Data-sized call of a distance helper: the (k,1,d) x (1,M,d) broadcast of kmeans.py:79 or
    any other broadcast (element by element) through yb_kmeans_dist."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = a.shape[-1]
    if b.shape[-1] != d:
        raise ValueError("operands could not be broadcast together")
    dev = _device()
    if a.ndim == 3 and b.ndim == 3 and a.shape[1] == 1 and b.shape[0] == 1:
        ta = torch.from_numpy(np.ascontiguousarray(a[:, 0, :])).to(dev)
        tb = torch.from_numpy(np.ascontiguousarray(b[0, :, :])).to(dev)
        return engine.kmeans_dist(ta, tb, what, outer=True).cpu().numpy()
    lead = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    ab = np.ascontiguousarray(np.broadcast_to(a, lead + (d,))).reshape(-1, d)
    bb = np.ascontiguousarray(np.broadcast_to(b, lead + (d,))).reshape(-1, d)
    out = engine.kmeans_dist(torch.from_numpy(ab).to(dev), torch.from_numpy(bb).to(dev), what, outer=False)
    return out.cpu().numpy().reshape(lead)


def _big(*arrays):
    return max(np.size(a) for a in arrays) > _HOST_LIMIT


def iou(center_boxes, data_boxes):
    """Area ratio min/max (the reference's ``iou``; not a box overlap)."""
    if _big(center_boxes, data_boxes):
        return _on_device(0, center_boxes, data_boxes)
    ca = center_boxes[..., 0] * center_boxes[..., 1]
    da = data_boxes[..., 0] * data_boxes[..., 1]
    return np.minimum(ca, da) / np.maximum(ca, da)


def iou_dist(center_boxes, data_boxes):
    if _big(center_boxes, data_boxes):
        return _on_device(1, center_boxes, data_boxes)
    return 1 - iou(center_boxes, data_boxes)


def euclidean_dist(center_boxes, data_boxes):
    if _big(center_boxes, data_boxes) and np.shape(center_boxes)[-1] < 8:
        return _on_device(2, center_boxes, data_boxes)
    return np.sqrt(np.sum(np.square(center_boxes - data_boxes), axis=-1))


_KIND = {iou_dist: YB_DIST_IOU, euclidean_dist: YB_DIST_EUCLID}
_PROBE_A = np.array([[[0.31, 0.27]], [[0.052, 0.91]], [[0.66, 0.044]]])
_PROBE_B = np.array([[[0.12, 0.80], [0.45, 0.33], [0.0071, 0.019], [0.97, 0.62]]])


def _kind_of(dist_func):
    """Which kernel a distance function maps to.  This module's two functions by identity; any
    other callable (e.g. the reference's own ``utils.kmeans.iou_dist`` object) by behaviour on a
    probe: it must reproduce one of them bit for bit.  Anything else cannot run on the GPU."""
    if dist_func in _KIND:
        return _KIND[dist_func]
    try:
        got = np.asarray(dist_func(_PROBE_A, _PROBE_B), dtype=np.float64)
    except Exception as e:  # noqa: BLE001
        raise YoloB200Error(f"dist_func failed on the probe boxes: {e}") from e
    for fn, kind in _KIND.items():
        if got.shape == (3, 4) and np.array_equal(got, fn(_PROBE_A, _PROBE_B)):
            return kind
    raise YoloB200Error("dist_func is neither iou_dist nor euclidean_dist (of this module or of the reference): "
                        "arbitrary Python distance functions cannot run on the GPU, and there is no CPU fallback")


def _bcast(arr, process_group):
    """rank 0's values on every rank (random draws must agree: ADVICE r1)."""
    if process_group is None or dist_util.world(process_group)[0] == 1:
        return arr
    import torch.distributed as dist
    backend = dist.get_backend(process_group)
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not dist.group.WORLD else 0,
                   group=process_group)
    return t.cpu().numpy()


def kmeans(data, n_cluster, dist_func, stop_dist, max_iternum=10000, verbose=True,
           process_group=None, return_assignments=False, check_every=8, exchange="nccl"):
    """k-means with the reference's semantics; ``data`` is (num_samples, num_dims) as an ndarray
    or a float64 CUDA tensor (this rank's shard when ``process_group`` is given).
    ``check_every`` iterations are queued between two looks at the loop state (any value gives the
    same centres).  ``return_assignments``: also the assignment of every box against the returned
    (float64, pre-rounding) centres.  ``exchange``: how the k*(d+1) partial sums of a sharded
    iteration are all-reduced - "nccl" (an all-reduce on the same stream between the assignment
    launch and a small update launch) or "peer" (inside the assignment launch, over NVLink peer
    memory: one launch per iteration; ranks of one node, NCCL group for the handle exchange)."""
    kind = _kind_of(dist_func)
    own_dist = iou_dist if kind == YB_DIST_IOU else euclidean_dist
    if torch.is_tensor(data):
        dev_data = data.to(torch.float64).contiguous()
        if not dev_data.is_cuda:
            dev_data = dev_data.to(_device())
    else:
        dev_data = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64))).to(_device())
    n_dim = dev_data.shape[-1]
    dev_data = dev_data.reshape(-1, n_dim)
    dev = dev_data.device
    sharded = process_group is not None and dist_util.world(process_group)[0] > 1
    mm = engine.minmax(dev_data)
    if sharded:
        mm = dist_util.allreduce_minmax(mm, process_group)
    data_min, data_max = (np.float64(v) for v in mm.cpu().numpy())

    center = rand(n_cluster * n_dim).reshape((n_cluster, 1, n_dim)) * data_max
    center = center * (data_max - data_min) + data_min
    center = _bcast(center, process_group)

    with torch.cuda.device(dev):
        dev_center = torch.from_numpy(np.ascontiguousarray(center.reshape(n_cluster, n_dim))).to(dev)
        if exchange not in ("nccl", "peer"):
            raise ValueError(f"Invalid exchange: {exchange}")
        peer = process_group if (sharded and exchange == "peer") else None
        loop = engine.KMeansLloyd(dev_data, dev_center, kind, stop_dist, max_iternum, sharded=sharded, peer_group=peer)
        seen = 0          # updates whose loss has been reported
        check_every = max(1, min(int(check_every), engine.N.YB_KMEANS_HIST))
        while True:
            if sharded and peer is None:
                for _ in range(check_every):
                    loop.step()
                    dist_util.allreduce_sum(loop.packed, process_group)   # same stream, no host sync
                    loop.update()
            else:   # the whole iteration is one launch: the batch is one CUDA-graph replay
                loop.step_many(check_every)
            status, done, hist, sums, counts = loop.read_state()
            if verbose:
                for e in range(seen + 1, done + 1):
                    print(f"epoch {e:2d}: loss = {hist[(e - 1) % len(hist)]:.4f}")
            seen = done
            if status == 3:
                # an empty cluster: this update is the host's (numpy.random owns the re-draw, kmeans.py:89)
                center = dev_center.cpu().numpy().reshape(n_cluster, 1, n_dim)
                new_center = np.copy(center)
                draws = []
                for n in range(n_cluster):
                    if counts[n] > 0:
                        cluster = sums[n] / counts[n]
                    else:
                        cluster = rand(n_dim) * (data_max - data_min) + data_min
                        draws.append(n)
                    new_center[n, 0] = cluster
                if draws and sharded:
                    new_center = _bcast(new_center, process_group)
                loss = np.mean(own_dist(center, new_center))
                done += 1
                seen = done
                if verbose:
                    print(f"epoch {done:2d}: loss = {loss:.4f}")
                dev_center.copy_(torch.from_numpy(np.ascontiguousarray(new_center.reshape(n_cluster, n_dim))))
                status = 1 if loss < stop_dist else (2 if done + 1 > max_iternum else 0)
                loop.resume(status, done)
            if status == 4:
                raise YoloB200Error("k-means peer exchange timed out: a rank did not issue its iteration")
            if status != 0:
                break
        center = dev_center.cpu().numpy()
        assign = None
        if return_assignments:
            assign, _, _ = engine.kmeans_assign(dev_data, dev_center, kind, want_assign=True)
        if peer is not None:
            import torch.distributed as dist
            dist.barrier(group=process_group)
            loop.close()
    center = center.reshape((n_cluster, n_dim)).astype("float32")
    if return_assignments:
        return center, assign
    return center
