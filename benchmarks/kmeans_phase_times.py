"""Phase times inside kmeans_assign_kernel (%globaltimer stamps of every CTA: tables, streaming, CTA
reduction, last-CTA reduction, peer exchange, Lloyd update) and the per-iteration time of graph batches
at 6.25 M and 50 M boxes.  Needs  make -C tf2_yolo_b200/csrc clean all EXTRA=-DYB_KM_PROF  (the default
build has no stamps)."""
import sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from tf2_yolo_b200 import engine, synth, _native as N
from tf2_yolo_b200._native import YB_DIST_IOU
dev = torch.device("cuda", 0)
rng = np.random.default_rng(4)
k = 9; n = 6_250_000
data = torch.from_numpy(synth.make_kmeans_boxes(rng, n, k)).to(dev)
c0 = np.sort(rng.uniform(0.02, 0.8, (k, 2)), axis=0)
if not hasattr(N.lib, "yb_debug_km_prof"):
    raise SystemExit(__doc__)
f = N.lib.yb_debug_km_prof; f.argtypes = [C.c_void_p, C.c_int]
for peer in (None, "self"):
    loop = engine.KMeansLloyd(data, torch.from_numpy(c0).to(dev), YB_DIST_IOU, 0.0, 1 << 40, sharded=peer is not None, peer_group=peer)
    loop.step_many(8); loop.step_many(8); loop.step_many(8)
    torch.cuda.synchronize()
    buf = np.zeros(1024 * 8, dtype=np.uint64)
    assert f(buf.ctypes.data, buf.size) == 0
    p = buf.reshape(1024, 8).astype(np.int64)
    p = p[p[:, 0] > 0]
    t0 = p[:, 0].min()
    last = np.argmax(p[:, 5])      # the last CTA of the final launch has the largest stamp 5
    print("peer", peer, "ctas", len(p))
    names = ["entry", "tables+issue", "stream_done", "sync", "partials+ticket", "last:reduce", "last:exchange", "last:update"]
    for i, nm in enumerate(names[:5]):
        print(f"  {nm:16s} min {p[:, i].min() - t0:7d} mean {p[:, i].mean() - t0:9.0f} max {p[:, i].max() - t0:7d}")
    for i in (5, 6, 7):
        print(f"  {names[i]:16s} {p[last, i] - t0:7d}")
    loop.close()
for n2 in (6_250_000, 50_000_000):
    if n2 != n:
        del data
        data = torch.from_numpy(synth.make_kmeans_boxes(rng, n2, k)).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for peer in (None, "self"):
        loop = engine.KMeansLloyd(data, torch.from_numpy(c0).to(dev), YB_DIST_IOU, 0.0, 1 << 40, sharded=peer is not None, peer_group=peer)
        loop.step_many(8); loop.step_many(8)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            loop.step_many(8)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"boxes": n2, "peer": peer, "graph_us": e0.elapsed_time(e1) / 40 * 1e3, "status": loop.read_state()[0]}), flush=True)
        loop.close()
