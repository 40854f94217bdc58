"""AP table of N synthetic v4-608 images: the CUDA path against the CPU oracle, in two halves so that
the (slow) oracle does not have to run on the GPU box:

    python benchmarks/map_subsample_check.py gpu --images 500 --out gpurun_out/map500_gpu.npz    # on the B200
    python benchmarks/map_subsample_check.py cpu --images 500 --ref gpurun_out/map500_gpu.npz \
           --json profiles/r2/map500_oracle_check.json                                           # anywhere

Both halves build the same images from the same seed (synth.make_config) and use the settings of
benchmarks/bench_multi.py map5."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KW = dict(conf_threshold=0.05, nms_mode=1, nms_threshold=0.5, max_per_img=100, version=4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("side", choices=["gpu", "cpu"])
    ap.add_argument("--images", type=int, default=500)
    ap.add_argument("--out", default="gpurun_out/map_subsample_gpu.npz")
    ap.add_argument("--ref", default="gpurun_out/map_subsample_gpu.npz")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from tf2_yolo_b200 import synth
    cfg = synth.make_config("v4-608", batch=a.images, seed=50)
    names = [str(i) for i in range(80)]
    y_true = cfg["y_trues"][-1].astype(np.float64)
    t0 = time.perf_counter()
    if a.side == "gpu":
        from tf2_yolo_b200.utils import measurement as meas
        pr = meas.PRfunc(y_true, *cfg["y_preds"], class_names=names, **KW)
        tab = pr.get_map()
        dt = time.perf_counter() - t0
        np.savez(a.out, ap=tab["ap"].values.astype(np.float64), images=a.images, seconds=dt,
                 checksum=np.float64(sum(float(p.sum()) for p in cfg["y_preds"])))
        print(json.dumps({"side": "gpu", "images": a.images, "seconds": dt, "mAP": float(tab["ap"].iloc[-1])}))
        return
    from oracle import measurement as om
    ref = np.load(a.ref)
    if int(ref["images"]) != a.images:
        raise SystemExit("the GPU half was run on a different number of images")
    if float(ref["checksum"]) != sum(float(p.sum()) for p in cfg["y_preds"]):
        raise SystemExit("the two halves did not build the same synthetic images")
    got = om.PRfunc(y_true, *cfg["y_preds"], class_names=names, **KW).get_ap("voc2012")
    dt = time.perf_counter() - t0
    got = np.asarray(got, dtype=np.float64)
    res = {"images": a.images, "classes": 80, "seconds_cpu_oracle": dt, "seconds_gpu": float(ref["seconds"]),
           "ap_table_bit_identical": bool(np.array_equal(got, ref["ap"])),
           "max_abs_diff": float(np.max(np.abs(got - ref["ap"]))), "mAP_voc2012": float(ref["ap"][-1]),
           "settings": KW}
    print(json.dumps(res))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
