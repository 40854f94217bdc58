"""Validate the loss oracle's TensorFlow->torch shim against REAL TensorFlow (run OUTSIDE the
build container, where `import tensorflow` works).

    YB_REFERENCE_ROOT=/path/to/tf2_YOLO python -m oracle.validate_with_tensorflow

The committed fixtures tests/golden/loss.npz and metrics.npz were produced by executing the
reference's `yolov*/losses/loss.py` and `yolov*/metrics/yolo_metrics.py` verbatim over a 16-symbol
TF->torch shim (oracle/refexec.py), because TensorFlow cannot be installed in the B200 image.
This script runs the SAME unmodified closures under real TensorFlow with `tf.GradientTape` on the
fixtures' inputs and compares loss and dL/dy_pred with the stored values:
float64 inputs: expected agreement ~1e-12 relative (what the restatement achieves against the
shim); float32 inputs: the north-star tolerances (loss 1e-5, gradient 1e-4).  Exit code 0 = the
shim, and therefore every parity claim pinned to these fixtures, holds against TensorFlow itself.
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YB_REFERENCE_ROOT", "/root/reference")
PKG = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import tensorflow as tf
    sys.path.insert(0, REF)
    z = np.load(os.path.join(ROOT, "tests", "golden", "loss.npz"), allow_pickle=False)
    worst = dict(loss64=0.0, grad64=0.0, loss32=0.0, grad32=0.0)
    for name in z["names"]:
        name = str(name)
        meta = json.loads(str(z[name + "/meta"]))
        kw = dict(meta["kwargs"])
        if kw.pop("binary_weight_is_array", False):
            kw["binary_weight"] = np.asarray(kw["binary_weight"], dtype=np.float64)
        mod = load(os.path.join(REF, PKG[meta["version"]], "losses", "loss.py"), "ref_loss_" + name)
        for dt, tag in ((tf.float64, "64"), (tf.float32, "32")):
            k = dict(kw)
            if k.get("anchors") is not None:
                k["anchors"] = np.asarray(k["anchors"], dtype=dt.as_numpy_dtype)
            if isinstance(k.get("binary_weight"), np.ndarray):
                k["binary_weight"] = k["binary_weight"].astype(dt.as_numpy_dtype)
            fn = mod.wrap_yolo_loss(grid_shape=(meta["grid"],) * 2, bbox_num=meta["B"], class_num=meta["C"], **k)
            yt = tf.constant(z[name + "/y_true"], dtype=dt)
            yp = tf.Variable(tf.constant(z[name + "/y_pred"], dtype=dt))
            with tf.GradientTape() as tape:
                loss = fn(yt, yp)
            grad = tape.gradient(loss, yp).numpy().astype(np.float64)
            l_ref, g_ref = float(z[name + "/loss"][0]), z[name + "/grad"]
            el = abs(float(np.asarray(loss).reshape(-1)[0]) - l_ref) / abs(l_ref)
            eg = np.abs(grad - g_ref).max() / np.abs(g_ref).max()
            worst["loss" + tag] = max(worst["loss" + tag], el)
            worst["grad" + tag] = max(worst["grad" + tag], eg)
            print(f"{name:16s} {tag}: loss rel {el:.2e}  grad rel(max) {eg:.2e}")
    print("worst:", worst)
    ok = worst["loss64"] <= 1e-10 and worst["grad64"] <= 1e-10 and worst["loss32"] <= 1e-5 and worst["grad32"] <= 1e-4
    # in-training metrics
    zm = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"), allow_pickle=False)
    for name in zm["names"]:
        name = str(name)
        meta = json.loads(str(zm[name + "/meta"]))
        pkg = PKG[meta["version"]]
        mm = importlib.import_module(f"{pkg}.metrics.yolo_metrics")
        g, B, Cn = (meta["grid"],) * 2, meta["B"], meta["C"]
        fns = [mm.wrap_obj_acc(g, B, Cn), mm.wrap_mean_iou(g, B, Cn),
               mm.wrap_class_acc(g, Cn) if meta["version"] == 1 else mm.wrap_class_acc(g, B, Cn),
               mm.wrap_recall(g, B, Cn, iou_threshold=meta["thr"])]
        yt = tf.constant(zm[name + "/y_true"], dtype=tf.float64)
        yp = tf.constant(zm[name + "/y_pred"], dtype=tf.float64)
        got = np.array([float(tf.reduce_mean(tf.cast(f(yt, yp), tf.float64))) for f in fns])
        e = np.abs(got - zm[name + "/metrics"]).max()
        print(f"{name:16s} metrics abs err {e:.2e}")
        ok = ok and e <= 1e-12
    print("SHIM VALIDATED AGAINST TENSORFLOW" if ok else "MISMATCH: the shim does not reproduce TensorFlow")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
