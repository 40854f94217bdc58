"""Phase times inside decode_nms_image_kernel (one CTA per image), read from %globaltimer stamps.
Needs a library built with the stamps compiled in:  make -C tf2_yolo_b200/csrc clean all EXTRA=-DYB_FUSED_PROF
(the default build has none).  Prints mean / max / min nanoseconds per phase over the 128 CTAs of one
config-2 step."""
import ctypes as C, numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from tf2_yolo_b200 import engine, synth, _native as N
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
dev = torch.device("cuda", 0)
batch = 128
cfg = synth.make_config("v4-608", batch=batch, seed=2, rank=0)
B, Cn = cfg["bbox_num"], cfg["class_num"]
fns = [wrap_yolo_loss((S, S), B, Cn, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1], wh_reg_weight=0.01, ignore_thresh=0.6) for si, S in enumerate(cfg["grids"])]
params = [f.params for f in fns]
dev_p = [torch.from_numpy(a).to(dev) for a in cfg["y_preds"]]
dev_t = [torch.from_numpy(a).to(dev) for a in cfg["y_trues"]]
dpreds = [torch.empty_like(a) for a in dev_p]
for it in range(6):
    engine.loss_decode_nms_fused(params, dev_t, dev_p, 0.5, 0.45, 2, rows_per_img_cap=1024, dpreds=dpreds)
torch.cuda.synchronize()
buf = np.zeros(128 * 16, dtype=np.uint64)
if not hasattr(N.lib, "yb_debug_fused_prof"):
    raise SystemExit(__doc__)
f = N.lib.yb_debug_fused_prof
f.argtypes = [C.c_void_p, C.c_int]
rc = f(buf.ctypes.data, buf.size)
assert rc == 0, rc
p = buf.reshape(128, 16).astype(np.int64)
t0 = p[:, 0].min()
names = ["entry", "ticket", "hist", "cstart", "slots", "rows", "pairs", "order+scan+rank", "barrier", "prefix", "lookback", "outcopy"]
rel = p[:, :12] - t0
print("kernel span ns:", rel[:, 11].max())
print("start spread ns: max", rel[:, 0].max())
d = np.diff(rel, axis=1)
for i, n in enumerate(names[1:]):
    print(f"{n:12s} mean {d[:, i].mean():8.0f} max {d[:, i].max():8.0f} min {d[:, i].min():8.0f}")
print("end time: mean", rel[:, 11].mean(), "max", rel[:, 11].max())
print("time at greedy_all done: mean", rel[:, 8].mean(), "max", rel[:, 8].max())
print("n_rows mean/max", p[:, 12].mean(), p[:, 12].max(), "max class count mean/max", p[:, 13].mean(), p[:, 13].max())
worst = np.argsort(rel[:, 8])[-5:]
for w in worst:
    print("cta", w, "img", p[w, 14], "rows", p[w, 12], "maxc", p[w, 13], "phases", d[w].tolist())
