"""GPU parity: box lists -> label grids (yb_encode_labels) against the reference reader's fixtures
(utils/tools.py:179-209 + down2xlabel :342-367) and the oracle restatement on fresh inputs.
Bit-exact: float64 outputs equal the reference's arrays, float32 outputs equal their cast."""
import json

import numpy as np
import pytest
import torch

from oracle import tools as ot
from tf2_yolo_b200 import engine
from tf2_yolo_b200.utils import tools as yt

pytestmark = pytest.mark.gpu


def test_reference_fixtures(golden):
    z = golden("encode")
    for name in z["names"]:
        name = str(name)
        meta = json.loads(str(z[name + "/meta"]))
        L = meta["levels"]
        for dt in (np.float64, np.float32):
            got = yt.encode_labels(z[name + "/boxes"], z[name + "/offsets"], meta["size"], meta["grid"], meta["C"],
                                   n_levels=L, dtype=dt)
            assert len(got) == L
            for l in range(L):
                ref = z[name + f"/level{l}"].astype(dt)
                assert got[L - 1 - l].dtype == dt and np.array_equal(got[L - 1 - l], ref), (name, l, dt)


def random_boxes(rng, n_img, size, C, mean, spill=0.05):
    H, W = size
    counts = rng.poisson(mean, n_img)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nb = int(off[-1])
    cx, cy = rng.uniform(-spill * W, (1 + spill) * W, nb), rng.uniform(-spill * H, (1 + spill) * H, nb)
    w, h = rng.uniform(0, 0.6 * W, nb), rng.uniform(0, 0.6 * H, nb)
    return np.column_stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, rng.integers(0, C, nb)]), off


@pytest.mark.parametrize("size,grid,C,levels,mean,n_img", [
    ((608, 608), (76, 76), 80, 3, 10.0, 16),     # config 3 labels
    ((416, 416), (52, 52), 80, 3, 10.0, 8),      # config 2
    ((416, 416), (13, 13), 20, 1, 6.0, 8),       # config 1
    ((100, 90), (8, 12), 5, 3, 40.0, 7),         # many boxes per cell, non-integer cell sizes
    ((333, 517), (16, 24), 3, 4, 300.0, 3),      # hundreds of boxes per image, 4 levels
    ((64, 64), (2, 2), 1, 2, 3.0, 300),          # many tiny images
])
def test_vs_oracle(size, grid, C, levels, mean, n_img):
    rng = np.random.default_rng(hash((size, grid, C)) & 0xffff)
    boxes, off = random_boxes(rng, n_img, size, C, mean)
    ref = ot.encode_label_pyramid(boxes, off, size, grid, C, levels)
    got = yt.encode_labels(boxes, off, size, grid, C, n_levels=levels)
    for l in range(levels):
        assert np.array_equal(got[l], ref[l]), l
    got32 = yt.encode_labels(boxes, off, size, grid, C, n_levels=levels, dtype=np.float32)
    for l in range(levels):
        assert np.array_equal(got32[l], ref[l].astype(np.float32)), l


def test_empty_and_error_cases():
    # no images, images without boxes
    assert yt.encode_labels(np.zeros((0, 5)), [0], (64, 64), (4, 4), 3)[0].shape == (0, 4, 4, 8)
    out = yt.encode_labels(np.zeros((0, 5)), [0, 0, 0], (64, 64), (4, 4), 3, n_levels=2)
    assert out[0].shape == (2, 2, 2, 8) and out[1].shape == (2, 4, 4, 8)
    assert not out[0].any() and not out[1].any()
    # the reference raises on these; so does the mirror (nothing is silently encoded)
    with pytest.raises(ValueError):
        yt.encode_labels([[np.nan, 0, 10, 10, 0]], [0, 1], (64, 64), (4, 4), 3)
    with pytest.raises(IndexError):
        yt.encode_labels([[0, 0, 10, 10, 3]], [0, 1], (64, 64), (4, 4), 3)       # class 3 of 3
    with pytest.raises(IndexError):
        yt.encode_labels([[-400, 0, -390, 10, 0]], [0, 1], (64, 64), (4, 4), 3)   # column -25 of 4
    # odd grids cannot be halved (the reference indexes out of bounds)
    with pytest.raises(Exception):
        yt.encode_labels([[0, 0, 10, 10, 0]], [0, 1], (64, 64), (5, 5), 3, n_levels=2)


def test_device_resident_pyramid_feeds_the_loss():
    """The float32 pyramid written on the device is what the loss reads: same loss as with the
    host-encoded labels (y_true never crosses PCIe)."""
    from tf2_yolo_b200 import synth
    from tf2_yolo_b200.grid_loss import fused_losses
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    cfg = synth.make_config("v4-608", batch=4, seed=3)
    boxes, off = synth.boxes_from_labels(cfg["y_trues"][-1], (608, 608))
    dev = torch.device("cuda")
    labs, n_bad = engine.encode_labels(torch.from_numpy(boxes).to(dev), torch.from_numpy(off).to(dev), (608, 608),
                                       (76, 76), 80, n_levels=3, max_boxes_per_img=int(np.diff(off).max()))
    ref = ot.encode_label_pyramid(boxes, off, (608, 608), (76, 76), 80, 3, dtype=np.float32)
    assert int(n_bad.item()) == 0
    for a, b in zip(labs, ref):
        assert np.array_equal(a.cpu().numpy(), b)
    B, C = 3, 80
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yps = [torch.from_numpy(a).to(dev) for a in cfg["y_preds"]]
    l_dev, _, _ = fused_losses(fns, labs, yps)
    l_host, _, _ = fused_losses(fns, [torch.from_numpy(a).to(dev) for a in ref], yps)
    assert torch.equal(l_dev, l_host)
