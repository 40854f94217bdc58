"""GPU parity: PRfunc / get_map / create_score_mat mirrors vs the fixtures the reference produced."""
import json

import numpy as np
import pytest

from oracle import measurement as om
from tf2_yolo_b200 import synth
from tf2_yolo_b200.utils import measurement as meas

pytestmark = pytest.mark.gpu


def test_prfunc_matches_reference_fixtures(golden):
    z = golden("map")
    C = int(z["class_num"])
    names = [f"c{i}" for i in range(C)]
    preds = [z["pred0"], z["pred1"]]
    for vn in "abc":
        kw = json.loads(str(z[f"{vn}/kwargs"]))
        pr = meas.PRfunc(z["y_true"], *preds, class_names=names, version=4, **kw)
        for k in range(C):
            assert np.array_equal(pr.precisions[k], z[f"{vn}/precisions{k}"]), (vn, k)
            assert np.array_equal(pr.recalls[k], z[f"{vn}/recalls{k}"]), (vn, k)
        for mode in ("voc2007", "voc2012", "area", "smootharea"):
            tab = pr.get_map(mode)
            assert list(tab.index) == names + ["mAP"] and list(tab.columns) == ["ap"]
            assert np.array_equal(tab["ap"].to_numpy(dtype=np.float64), z[f"{vn}/ap_{mode}"]), (vn, mode)
        got = np.array([[pr(r, k) for r in (0.0, 0.3, 0.6, 0.9)] for k in range(C)], dtype=np.float64)
        assert np.array_equal(got, z[f"{vn}/call"])
    with pytest.raises(IndexError):
        pr(0.5, class_idx=C)


def test_score_mat_matches_reference_fixtures(golden):
    z = golden("map")
    C = int(z["class_num"])
    names = [f"c{i}" for i in range(C)]
    preds = [z["pred0"], z["pred1"]]
    for vn in ("s0", "s1", "s2"):
        kw = json.loads(str(z[f"{vn}/kwargs"]))
        tab = meas.create_score_mat(z["y_true"], *preds, class_names=names, version=4, **kw)
        assert list(tab.columns) == ["precision", "recall", "F1-score", "gts", "dets"]
        assert list(tab.index) == names
        got = tab.to_numpy(dtype=np.float64)
        assert np.array_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(z[f"{vn}/table"], nan=-7.0)), vn


def test_prfunc_larger_vs_oracle():
    """More images / classes than the fixture, truncation by max_per_img, DIoU-NMS."""
    rng = np.random.default_rng(77)
    grids, B, C, n_img = [8, 16], 3, 6, 40
    y_trues = synth.make_labels(rng, n_img, grids, C, synth.ANCHORS_V4[3:9], mean_boxes=6.0)
    y_preds = synth.make_head_outputs(rng, y_trues, grids, B, C, synth.ANCHORS_V4[3:9], det_per_gt=20,
                                      stray_frac=0.05)
    names = [f"c{i}" for i in range(C)]
    yt = y_trues[-1].astype(np.float64)
    for kw in (dict(conf_threshold=0.05, nms_mode=1, max_per_img=100),
               dict(conf_threshold=0.2, nms_mode=3, nms_threshold=0.45, max_per_img=4, precision_mode=0),
               dict(conf_threshold=0.2, nms_mode=0, max_per_img=None, precision_mode=1),
               dict(conf_threshold=0.3, nms_mode=2, nms_threshold=0.45, nms_sigma=0.5, max_per_img=100)):
        ref = om.PRfunc(yt, *y_preds, class_names=names, version=4, **kw)
        got = meas.PRfunc(yt, *y_preds, class_names=names, version=4, **kw)
        for k in range(C):
            assert np.array_equal(got.precisions[k], ref.precisions[k]), (kw, k)
            assert np.array_equal(got.recalls[k], ref.recalls[k]), (kw, k)
        assert np.allclose(got.get_map("voc2012")["ap"].to_numpy(), ref.get_ap("voc2012"), rtol=0, atol=1e-15)
    t = meas.create_score_mat(yt, *y_preds, class_names=names, conf_threshold=0.3, nms_mode=1, version=4)
    r = om.score_table(yt, *y_preds, class_names=names, conf_threshold=0.3, nms_mode=1, version=4)
    for col in ("precision", "recall", "F1-score", "gts", "dets"):
        assert np.array_equal(np.nan_to_num(t[col].to_numpy(dtype=np.float64), nan=-7),
                              np.nan_to_num(np.asarray(r[col], dtype=np.float64), nan=-7)), col
