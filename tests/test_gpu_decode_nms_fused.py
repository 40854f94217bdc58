"""GPU parity: the one-launch-per-batch decode + NMS (yb_decode_nms, one CTA per image) and the
two-launch train-and-evaluate step (yb_loss_decode_nms_fused) against the six-launch chain
(yb_decode + yb_nms), the oracle and the reference fixtures.  Bit-exact rows, order and offsets."""
import numpy as np
import pytest
import torch

from oracle import tools as ot
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200.grid_loss import fused_losses

pytestmark = pytest.mark.gpu


def chain(preds, C, thr, version, nms_thr, mode):
    rows, offs = engine.decode_batch_exact(preds, C, thr, version)
    g = engine.nms_batch(rows, offs, C, nms_thr, mode)
    n = int(g["out_offsets"][-1].item())
    return g["out_rows"][:n].cpu().numpy(), g["out_offsets"].cpu().numpy()


def fused(preds, C, thr, version, nms_thr, mode, cap=1024):
    r = engine.decode_nms_batch(preds, C, thr, version, nms_thr, mode, rows_per_img_cap=cap)
    offs = r["out_offsets"].cpu().numpy()
    return r["out_rows"][:offs[-1]].cpu().numpy(), offs, int(r["n_overflow"].item())


def test_reference_fixtures(golden):
    z = golden("decode_nms")
    preds = [torch.from_numpy(z["m/pred0"]).cuda(), torch.from_numpy(z["m/pred1"]).cuda()]
    for thr in (0.5, 0.3):
        for mode in (1, 2):
            rows, offs, ovf = fused(preds, 5, thr, 4, 0.45, mode)
            assert ovf == 0
            for i in range(3):
                key = f"m/nms{mode}_t{thr}_i{i}"
                ref = z[key] if key in z.files else np.zeros((0, 7))
                assert np.array_equal(rows[offs[i]:offs[i + 1]], ref.reshape(-1, 7)), (thr, mode, i)
    v2 = torch.from_numpy(z["v2/pred"]).cuda()
    rows, offs, ovf = fused([v2], 20, 0.4, 2, 0.45, 1)
    for i in range(2):
        assert np.array_equal(rows[offs[i]:offs[i + 1]], z[f"v2/nms1_i{i}"].reshape(-1, 7))
    # v1 layout (shared class scores per cell)
    v1 = torch.from_numpy(z["v1/pred"]).cuda()
    rows, offs, ovf = fused([v1], 6, 0.5, 1, 0.45, 1)
    for i in range(2):
        ref = ot.nms(z[f"v1/rows_i{i}"].reshape(-1, 7), 6, 0.45, 1)
        assert np.array_equal(rows[offs[i]:offs[i + 1]], ref)


@pytest.mark.parametrize("name,version,batch,thr", [("v4-608", 4, 12, 0.5), ("v3-416", 3, 9, 0.5), ("v2-416", 2, 8, 0.4),
                                                    ("v4-608", 4, 3, 0.3)])
@pytest.mark.parametrize("mode", [1, 2])
def test_equals_the_six_launch_chain(name, version, batch, thr, mode):
    cfg = synth.make_config(name, batch=batch, seed=13 + version)
    preds = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    C = cfg["class_num"]
    ref_rows, ref_offs = chain(preds, C, thr, version, 0.45, mode)
    rows, offs, ovf = fused(preds, C, thr, version, 0.45, mode, cap=2048)
    assert ovf == 0
    assert np.array_equal(offs, ref_offs)
    assert np.array_equal(rows, ref_rows)
    # and the oracle, image 0
    per = [p[0] for p in cfg["y_preds"]]
    o = ot.nms(ot.decode(*per, class_num=C, threshold=thr, version=version).reshape(-1, 7), C, 0.45, mode)
    assert np.array_equal(rows[offs[0]:offs[1]], o)


def test_ties_duplicates_and_big_segments():
    """Equal confidences, exact duplicates and one class holding hundreds of rows (a single warp
    sweeps any segment that fits the image's row cap)."""
    rng = np.random.default_rng(3)
    S, B, C = 8, 3, 4
    p = rng.uniform(0.02, 0.2, (3, S, S, B, 5 + C)).astype(np.float32)
    p[..., 2:4] = rng.uniform(0.2, 0.6, (3, S, S, B, 2))
    p[..., 4] = 0.9
    p[0, :, :, :, 5] = 0.8                       # image 0: every box a hit of class 0 (192 rows, ties everywhere)
    p[1, :4, :, :, 5 + 1] = rng.uniform(0.6, 0.9, (4, S, B))
    p[1, 2, 3] = p[1, 2, 2]                      # duplicated cell contents
    p[2, :, :, 0, 5:] = 0.7                      # image 2: 4 classes x 64 rows
    t = torch.from_numpy(p.reshape(3, S, S, B * (5 + C))).cuda()
    for mode in (1, 2):
        ref_rows, ref_offs = chain([t], C, 0.5, 3, 0.3, mode)
        rows, offs, ovf = fused([t], C, 0.5, 3, 0.3, mode, cap=512)
        assert ovf == 0 and np.array_equal(offs, ref_offs) and np.array_equal(rows, ref_rows)
        o = ot.nms(ot.decode(p[0].reshape(S, S, -1), class_num=C, threshold=0.5, version=3).reshape(-1, 7), C, 0.3, mode)
        assert np.array_equal(rows[offs[0]:offs[1]], o)


def test_class_sizes_around_the_mask_width():
    """Classes of exactly 1, 32, 33, 64, 65, 128 (pair masks) and 129 rows (warp sweep) in one image,
    with heavy overlap and ties."""
    rng = np.random.default_rng(11)
    S, C = 24, 7
    sizes = [1, 32, 33, 64, 65, 128, 129]
    p = np.zeros((2, S * S, 5 + C), dtype=np.float32)
    for img in range(2):
        cells = rng.permutation(S * S)[:sum(sizes)]
        at = 0
        for c, n in enumerate(sizes):
            sel = cells[at:at + n]
            at += n
            p[img, sel, 0:2] = rng.uniform(0.0, 1.0, (n, 2))
            p[img, sel, 2:4] = rng.uniform(0.15, 0.5, (n, 2))
            p[img, sel, 4] = np.round(rng.uniform(0.6, 1.0, n), 1 if img else 3)    # image 1: many equal confidences
            p[img, sel, 5 + c] = 0.9
    t = torch.from_numpy(p.reshape(2, S, S, 5 + C)).cuda()
    for mode in (1, 2):
        for nms_thr in (0.1, 0.45):
            ref_rows, ref_offs = chain([t], C, 0.5, 3, nms_thr, mode)
            rows, offs, ovf = fused([t], C, 0.5, 3, nms_thr, mode, cap=512)
            assert ovf == 0 and np.array_equal(offs, ref_offs) and np.array_equal(rows, ref_rows)
            assert 0 < offs[-1] < 2 * sum(sizes)
    o = ot.nms(ot.decode(p[1].reshape(S, S, -1), class_num=C, threshold=0.5, version=3).reshape(-1, 7), C, 0.45, 2)
    assert np.array_equal(rows[offs[1]:offs[2]], o)


def test_overflow_is_reported_and_the_checked_form_falls_back():
    cfg = synth.make_config("v4-608", batch=4, seed=21)
    preds = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    ref_rows, ref_offs = chain(preds, 80, 0.5, 4, 0.45, 2)
    per_img = np.diff(engine.decode_batch_exact(preds, 80, 0.5, 4)[1].cpu().numpy())
    cap = int(max(32, np.sort(per_img)[1]))          # at least two images do not fit
    rows, offs, ovf = fused(preds, 80, 0.5, 4, 0.45, 2, cap=cap)
    assert ovf == int((per_img > cap).sum()) and ovf >= 1
    for i in range(4):                               # images that fit are complete, the others empty
        got = rows[offs[i]:offs[i + 1]]
        want = ref_rows[ref_offs[i]:ref_offs[i + 1]]
        assert np.array_equal(got, want) if per_img[i] <= cap else got.shape[0] == 0
    r, o = engine.decode_nms_batch_exact(preds, 80, 0.5, 4, 0.45, 2, rows_per_img_cap=cap)
    assert np.array_equal(r.cpu().numpy(), ref_rows) and np.array_equal(o.cpu().numpy(), ref_offs)


def test_empty_inputs():
    z = torch.zeros((2, 4, 4, 3 * 9), device="cuda")
    rows, offs, ovf = fused([z], 4, 0.5, 3, 0.45, 1, cap=64)
    assert rows.shape[0] == 0 and list(offs) == [0, 0, 0] and ovf == 0
    r = engine.decode_nms_batch([torch.zeros((0, 4, 4, 27), device="cuda")], 4, 0.5, 3)
    torch.cuda.synchronize()
    assert int(r["out_offsets"][0].item()) == 0


def test_two_launch_step_equals_the_separate_calls():
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    cfg = synth.make_config("v4-608", batch=6, seed=5)
    B, C = 3, 80
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    loss0, d0, _ = fused_losses(fns, yts, yps)
    ref_rows, ref_offs = chain(yps, C, 0.5, 4, 0.45, 2)
    loss, d, _, r = engine.loss_decode_nms_fused([f.params for f in fns], yts, yps, 0.5, 0.45, 2, rows_per_img_cap=1024)
    assert torch.equal(loss, loss0) and all(torch.equal(a, b) for a, b in zip(d, d0))
    offs = r["out_offsets"].cpu().numpy()
    assert int(r["n_overflow"].item()) == 0
    assert np.array_equal(offs, ref_offs) and np.array_equal(r["out_rows"][:offs[-1]].cpu().numpy(), ref_rows)
    # determinism: the same bits on every run
    for _ in range(3):
        loss2, _, _, r2 = engine.loss_decode_nms_fused([f.params for f in fns], yts, yps, 0.5, 0.45, 2)
        assert torch.equal(loss2, loss0) and torch.equal(r2["out_rows"][:offs[-1]], r["out_rows"][:offs[-1]])


def test_checked_form_falls_back_for_shapes_the_per_image_kernel_does_not_take():
    """class_num > 256 and float64 grids go through the general chain (same result, no error)."""
    rng = np.random.default_rng(5)
    p = rng.uniform(0.3, 1.0, (2, 3, 3, 1 * (5 + 300))).astype(np.float32)      # 300 classes
    t = torch.from_numpy(p).cuda()
    ref_rows, ref_offs = chain([t], 300, 0.9, 3, 0.45, 1)
    r, o = engine.decode_nms_batch_exact([t], 300, 0.9, 3, 0.45, 1)
    assert np.array_equal(r.cpu().numpy(), ref_rows) and np.array_equal(o.cpu().numpy(), ref_offs)
    with pytest.raises(ValueError):
        engine.decode_nms_batch([t], 300, 0.9, 3)
    t64 = torch.from_numpy(p[..., :5 + 4].astype(np.float64)).cuda()
    r64, o64 = engine.decode_nms_batch_exact([t64], 4, 0.5, 3, 0.45, 1)
    a, b = chain([t64], 4, 0.5, 3, 0.45, 1)
    assert np.array_equal(r64.cpu().numpy(), a) and np.array_equal(o64.cpu().numpy(), b)


@pytest.mark.parametrize("graph", [False, True])
def test_static_step_replays_without_memsets(graph):
    """TrainEvalStep (yb_loss_decode_nms_fused_clean): workspaces zeroed once, the kernels leave them
    zeroed; every replay - also after the inputs were refilled in place and after an overflowing
    step - gives the bits of the eager two-launch step."""
    from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss
    B, C = 3, 80
    cfgs = [synth.make_config("v4-608", batch=5, seed=s) for s in (5, 6)]
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfgs[0]["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfgs[0]["grids"])]
    params = [f.params for f in fns]
    yts = [torch.from_numpy(a).cuda() for a in cfgs[0]["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfgs[0]["y_preds"]]
    step = engine.TrainEvalStep(params, yts, yps, 0.5, 0.45, 2, rows_per_img_cap=1024, graph=graph)
    assert (step.graph is not None) == graph
    for rep in range(5):
        cfg = cfgs[rep & 1]
        for dst, src in zip(yts + yps, cfg["y_trues"] + cfg["y_preds"]):
            dst.copy_(torch.from_numpy(src))
        loss, d, _, r = step.run()
        loss0, d0, _, r0 = engine.loss_decode_nms_fused(params, yts, yps, 0.5, 0.45, 2, rows_per_img_cap=1024)
        n = int(r0["out_offsets"][-1].item())
        assert n > 0 and int(r["n_overflow"].item()) == 0
        assert torch.equal(loss, loss0) and all(torch.equal(a, b) for a, b in zip(d, d0))
        assert torch.equal(r["out_offsets"], r0["out_offsets"]) and torch.equal(r["out_rows"][:n], r0["out_rows"][:n])
        torch.cuda.synchronize()
        # only the row buckets (never read before they are rewritten) may hold anything
        assert int(step.lws.count_nonzero().item()) == 0
        head = step._aligned(step.fws) - step.fws.data_ptr()
        assert int(step.fws[head:head + 768].count_nonzero().item()) == 0
    # a step whose images overflow the row capacity also leaves the control block clean
    small = engine.TrainEvalStep(params, yts, yps, 0.5, 0.45, 2, rows_per_img_cap=32, graph=graph)
    for _ in range(2):
        _, _, _, r = small.run()
        assert int(r["n_overflow"].item()) == 5 and int(r["out_offsets"][-1].item()) == 0
    torch.cuda.synchronize()
    head = small._aligned(small.fws) - small.fws.data_ptr()
    assert int(small.fws[head:head + 768].count_nonzero().item()) == 0
