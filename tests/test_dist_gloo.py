"""CPU, world_size 2 over gloo: the multi-GPU host logic.  The per-rank compute is
the oracle here (no GPU in this container); what is under test is the sharding contract and the
collectives in tf2_yolo_b200/dist.py that the CUDA path uses unchanged over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kmeans as okm
from oracle import losses as ol
from tf2_yolo_b200 import dist as ydist
from tf2_yolo_b200 import synth

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        out[rank] = fn(rank)
    finally:
        dist.destroy_process_group()


def _run(fn):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), fn, out), nprocs=WORLD, join=True)
    return [out[r] for r in range(WORLD)]


def _loss_rank(rank):
    cfg = synth.make_config("v3-416", batch=4, seed=6)
    S, B, C = 13, 3, 80
    spec = ol.GridLossSpec(version=3, grid_shape=(S, S), bbox_num=B, class_num=C, anchors=cfg["anchors"][:3])
    a, b = ydist.shard_range(4, rank, WORLD)
    yt, yp = cfg["y_trues"][0][a:b], cfg["y_preds"][0][a:b]
    local, grad, _ = ol.loss_and_grad(spec, yt, yp)
    n_local = b - a
    # the kernel divides by the GLOBAL batch (inv_batch = 1/N_global): rescale the oracle's local mean
    t = torch.tensor([local * n_local / 4.0], dtype=torch.float64)
    ydist.allreduce_sum(t)
    full, gfull, _ = ol.loss_and_grad(spec, cfg["y_trues"][0], cfg["y_preds"][0])
    return float(t), full, float(np.abs(grad * n_local / 4.0 - gfull[a:b]).max())


def test_sharded_loss_sums_to_global_batch_loss():
    res = _run(_loss_rank)
    for total, full, gerr in res:
        assert abs(total - full) <= 1e-12 * abs(full)
        assert gerr <= 1e-12


def _kmeans_rank(rank):
    rng = np.random.default_rng(1)
    data = synth.make_kmeans_boxes(rng, 5001, k=4)
    centers = np.random.default_rng(2).uniform(0.05, 0.6, (4, 2))
    a, b = ydist.shard_range(len(data), rank, WORLD)
    who = okm.assign(data[a:b], centers)
    sums = torch.from_numpy(np.stack([data[a:b][who == c].sum(axis=0) for c in range(4)]))
    counts = torch.from_numpy(np.bincount(who, minlength=4).astype(np.int64))
    gs, gc = ydist.allreduce_kmeans(sums, counts)
    mm = ydist.allreduce_minmax(torch.tensor([data[a:b].min(), data[a:b].max()]))
    whole = okm.assign(data, centers)
    want = np.stack([data[whole == c].sum(axis=0) for c in range(4)])
    return (np.allclose(gs.numpy(), want, rtol=1e-13), gc.tolist() == np.bincount(whole, minlength=4).tolist(),
            mm.tolist() == [data.min(), data.max()])


def test_kmeans_partial_sums_allreduce():
    for ok in _run(_kmeans_rank):
        assert all(ok)


def _gather_rank(rank):
    local_counts = torch.tensor([3, 0, 5], dtype=torch.int64) * (rank + 1)
    before, total = ydist.rank_offsets(local_counts)
    t = torch.arange(4 + 3 * rank, dtype=torch.float64) + 100 * rank
    g = ydist.gather_varlen(t)
    e = ydist.gather_varlen(torch.zeros(0, dtype=torch.int32) if rank == 0 else torch.ones(2, dtype=torch.int32))
    return before.tolist(), total.tolist(), g.tolist(), e.tolist()


def test_rank_offsets_and_varlen_gather():
    res = _run(_gather_rank)
    assert res[0][0] == [0, 0, 0] and res[1][0] == [3, 0, 5]
    for r in res:
        assert r[1] == [9, 0, 15]
        assert r[2] == [0, 1, 2, 3, 100, 101, 102, 103, 104, 105, 106]
        assert r[3] == [1, 1]


def _route_rank(rank):
    """Each rank holds records of 5 classes in two class-major chunks; class c must end on rank
    c % 2 with its records in (rank, chunk, local) order."""
    C = 5
    rng = np.random.default_rng(100 + rank)
    blocks, segs, base = [], [], 0
    for chunk in range(2):
        counts = rng.integers(0, 4, C)
        counts[rank] = 0                       # an empty class per rank
        co = np.concatenate([[0], np.cumsum(counts)])
        segs.append(co + base)
        for c in range(C):
            for j in range(counts[c]):
                blocks.append((c, 1000 * rank + 100 * chunk + 10 * c + j))
        base += co[-1]
    cls = torch.tensor([b[0] for b in blocks], dtype=torch.int32)
    tag = torch.tensor([b[1] for b in blocks], dtype=torch.int64)
    conf = tag.to(torch.float64) / 7.0
    flag = (tag % 2).to(torch.uint8)
    out = ydist.route_records_by_class((conf, tag, flag, cls), np.stack(segs), C)
    return [t.tolist() for t in out]


def test_records_are_routed_to_their_class_owner_in_image_order():
    res = _run(_route_rank)
    all_tags = {}
    for rank in range(WORLD):
        conf, tag, flag, cls, owned = res[rank]
        assert all(c % WORLD == rank for c in cls)                      # only classes this rank owns
        assert [t / 7.0 for t in tag] == conf and [t % 2 for t in tag] == flag   # the arrays travel together
        for c in set(cls):
            mine = [t for t, k in zip(tag, cls) if k == c]
            assert mine == sorted(mine), (rank, c)                       # (rank, chunk, local) order kept
            assert owned[c] == len(mine)
            all_tags[c] = mine
        assert all(owned[c] == 0 for c in range(5) if c % WORLD != rank)
    # nothing lost: every record of every rank arrived exactly once
    total = sum(len(v) for v in all_tags.values())
    sent = 0
    for rank in range(WORLD):
        rng = np.random.default_rng(100 + rank)
        for chunk in range(2):
            counts = rng.integers(0, 4, 5)
            counts[rank] = 0
            sent += int(counts.sum())
    assert total == sent
