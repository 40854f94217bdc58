"""NumPy restatement of the reference's anchor k-means (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils/kmeans.py: area_ratio :9-24 (``iou`` there is
min(area)/max(area), NOT box overlap), iou_dist :27-33, euclidean_dist :36-40,
the Lloyd loop :43-102 including its use of the GLOBAL numpy RNG (one
``rand(k*d)`` for the initial centres, one ``rand(d)`` per empty cluster in
ascending cluster order per iteration).

Parity pinning: checked against the unmodified reference executed in the build
container (oracle/refexec.py) and tests/golden/kmeans.npz.
"""
import numpy as np
from numpy.random import rand


def area_ratio(centers, boxes):
    ca = centers[..., 0] * centers[..., 1]
    ba = boxes[..., 0] * boxes[..., 1]
    return np.minimum(ca, ba) / np.maximum(ca, ba)


def iou_dist(centers, boxes):
    return 1 - area_ratio(centers, boxes)


def euclidean_dist(centers, boxes):
    return np.sqrt(np.sum(np.square(centers - boxes), axis=-1))


def assign(data, centers, dist_func=iou_dist):
    """argmin over clusters for every sample; data (M,d), centers (k,d)."""
    d = dist_func(np.asarray(centers)[:, None, :], np.asarray(data)[None, :, :])
    return np.argmin(d, axis=0)


def lloyd_step(data, centers, dist_func, lo, hi):
    """One iteration: (new_centers (k,d), assignments, shift)."""
    k, nd = centers.shape
    who = assign(data, centers, dist_func)
    new = centers.copy()
    for c in range(k):
        idx = np.nonzero(who == c)[0]
        new[c] = np.mean(data[idx], axis=0) if len(idx) else rand(nd) * (hi - lo) + lo
    shift = np.mean(dist_func(centers[:, None, :], new[:, None, :]))
    return new, who, shift


def kmeans(data, n_cluster, dist_func, stop_dist, max_iternum=10000, verbose=False, trace=None):
    data = np.asarray(data)
    nd = data.shape[-1]
    hi, lo = data.max(), data.min()
    centers = rand(n_cluster * nd).reshape(n_cluster, nd) * hi
    centers = centers * (hi - lo) + lo
    epoch = 1
    while True:
        centers, who, shift = lloyd_step(data, centers, dist_func, lo, hi)
        if trace is not None:
            trace.append((centers.copy(), who, float(shift)))
        if verbose:
            print(f"epoch {epoch:2d}: loss = {shift:.4f}")
        epoch += 1
        if shift < stop_dist or epoch > max_iternum:
            break
    return centers.astype("float32")
