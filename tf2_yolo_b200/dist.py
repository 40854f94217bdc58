"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on the B200 box; the
same code runs over gloo on CPU tensors in the tests).  The hot path shards by image / box with
no data-path collective; these helpers carry the only exchanges there are:

* loss      : all-reduce of the per-scale scalars (each rank divides by the GLOBAL batch)
* k-means   : all-reduce of k*(d+1) partial sums per Lloyd iteration, min/max once
* PR / mAP  : all-gather of per-class ground-truth counts (to offset gt ids); the variable-length
              (conf, gt_id, flag, class) records go to the rank that owns their class (class % world,
              all-to-all), so the sort + scan of the PR curves is split 1/world per rank
"""
import torch
import torch.distributed as dist


def world(group=None):
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def shard_range(n, rank, n_ranks):
    """Contiguous, rank-ordered split of n items (first n % n_ranks ranks get one extra)."""
    base, extra = divmod(n, n_ranks)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_sum(t, group=None):
    if world(group)[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_minmax(mm, group=None):
    """mm = tensor [min, max] -> global [min, max]."""
    if world(group)[0] > 1:
        lo, hi = mm[0:1].clone(), mm[1:2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        mm = torch.cat([lo, hi])
    return mm


def allreduce_kmeans(sums, counts, group=None):
    """Per-cluster partial sums (k,d) float64 and counts (k,) int64 -> global ones, one collective."""
    if world(group)[0] == 1:
        return sums, counts
    k, d = sums.shape
    packed = torch.cat([sums.reshape(-1), counts.to(torch.float64)])   # counts < 2^53: exact
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:k * d].reshape(k, d), packed[k * d:].round().to(torch.int64)


def rank_offsets(local_counts, group=None):
    """(sum of local_counts over earlier ranks, total over all ranks); int64 tensors."""
    n, r = world(group)
    if n == 1:
        return torch.zeros_like(local_counts), local_counts.clone()
    bufs = [torch.zeros_like(local_counts) for _ in range(n)]
    dist.all_gather(bufs, local_counts, group=group)
    before = torch.zeros_like(local_counts)
    for i in range(r):
        before += bufs[i]
    return before, torch.stack(bufs).sum(0)


def gather_varlen(t, group=None):
    """Concatenate 1-D tensors of different lengths from all ranks, in rank order."""
    n, _ = world(group)
    if n == 1:
        return t
    size = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(size) for _ in range(n)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(n)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def route_records_by_class(arrays, segments, class_num, group=None):
    """All-to-all of per-class records: class c goes to rank c % world.

    ``arrays``: this rank's record arrays (same length), laid out as chunks whose blocks are
    class-major; ``segments`` (n_chunks, C+1) int64 ndarray locates class c of chunk k at
    [segments[k, c], segments[k, c+1]).  Returns (arrays of the records this rank owns, ordered
    source rank -> class -> chunk, i.e. every class keeps the global image order of its records;
    class_counts (C,) int64 tensor: global record count of the owned classes, 0 elsewhere)."""
    n, r = world(group)
    dev = arrays[0].device
    C = class_num
    # send order: destination rank, then class, then chunk
    pieces = [[] for _ in arrays]
    send_counts = []
    per_class = torch.zeros(C, dtype=torch.int64)
    for d in range(n):
        cnt = 0
        for c in range(d, C, n):
            for k in range(segments.shape[0]):
                a, b = int(segments[k, c]), int(segments[k, c + 1])
                if b > a:
                    for i, arr in enumerate(arrays):
                        pieces[i].append(arr[a:b])
                    cnt += b - a
                    per_class[c] += b - a
        send_counts.append(cnt)
    send = [torch.cat(p) if p else arr[:0] for p, arr in zip(pieces, arrays)]
    per_class = per_class.to(dev)
    if n == 1:
        return (*send, per_class)
    dist.all_reduce(per_class, group=group)            # global records per class
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu()]
    out = []
    for t in send:
        buf = torch.empty(sum(recv_counts), dtype=t.dtype, device=dev)
        dist.all_to_all_single(buf, t.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts,
                               group=group)
        out.append(buf)
    owned = torch.zeros(C, dtype=torch.int64, device=dev)
    owned[r::n] = per_class[r::n]
    return (*out, owned)
