"""GPU parity: fused loss forward+gradient vs the fp64 oracle and the reference fixtures.
Tolerances (north_star): loss 1e-5 relative, gradient 1e-4 relative, fp32 kernel vs fp64 truth."""
import numpy as np
import pytest
import torch

from conftest import loss_cases
from oracle import losses as ol
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200.grid_loss import fused_losses

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def wrap(version):
    import importlib
    pkg = {1: "yolov1_5", 2: "yolov2", 3: "yolov3", 4: "yolov4"}[version]
    return importlib.import_module(f"tf2_yolo_b200.{pkg}.losses").wrap_yolo_loss


def check_grad(g, g_ref, name=""):
    g_ref = np.asarray(g_ref, dtype=np.float64)
    scale = np.abs(g_ref).max()
    err = np.abs(g.astype(np.float64) - g_ref)
    tol = GRAD_RTOL * np.abs(g_ref) + 1e-6 * scale
    bad = err > tol
    assert not bad.any(), f"{name}: {bad.sum()} gradient entries off, worst {err.max():.3e} (scale {scale:.3e})"
    # zero pattern is exact: non-responsible class scores get exactly 0
    assert np.array_equal(g_ref == 0, g == 0) or (np.abs(g[g_ref == 0]).max() <= 1e-12 * scale), name


def test_reference_fixtures(golden):
    npz = golden("loss")
    for name, meta, kw, yt, yp, loss_ref, grad_ref in loss_cases(npz):
        S = meta["grid"]
        kwargs = dict(kw)
        if meta["version"] == 1:
            fn = wrap(1)((S, S), meta["B"], meta["C"], **kwargs)
        else:
            fn = wrap(meta["version"])((S, S), meta["B"], meta["C"], **kwargs)
        loss, grad = fn.value_and_grad(yt, yp)
        assert loss.shape == ((1,) if isinstance(kw.get("binary_weight"), np.ndarray) else ())
        assert abs(float(loss.reshape(-1)[0]) - loss_ref[0]) <= LOSS_RTOL * abs(loss_ref[0]), name
        check_grad(grad, grad_ref, name)


@pytest.mark.parametrize("name,version,batch", [("v4-608", 4, 3), ("v3-416", 3, 4), ("v2-416", 2, 8)])
def test_baseline_configs_vs_oracle(name, version, batch):
    cfg = synth.make_config(name, batch=batch, seed=version)
    B, C = cfg["bbox_num"], cfg["class_num"]
    fns, specs = [], []
    for si, S in enumerate(cfg["grids"]):
        anc = cfg["anchors"][si * B:(si + 1) * B]
        kw = dict(anchors=anc)
        if version == 4:
            kw.update(loss_weight=[1, 5, 1], wh_reg_weight=0.01)
        else:
            kw.update(loss_weight=[1, 1, 5, 1])
        fns.append(wrap(version)((S, S), B, C, **kw))
        specs.append(ol.GridLossSpec(version=version, grid_shape=(S, S), bbox_num=B, class_num=C, **kw))
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    # fused: all scales in one launch
    loss, dpreds, terms = fused_losses(fns, yts, yps, want_terms=True)
    loss = loss.cpu().numpy()
    for si in range(len(fns)):
        l_ref, g_ref, t_ref = ol.loss_and_grad(specs[si], cfg["y_trues"][si], cfg["y_preds"][si])
        assert abs(loss[si] - l_ref) <= LOSS_RTOL * abs(l_ref), (name, si, loss[si], l_ref)
        assert abs(terms[si, 0].item() - l_ref) <= LOSS_RTOL * abs(l_ref)
        check_grad(dpreds[si].cpu().numpy(), g_ref, f"{name}/s{si}")
        # one scale per launch (what Keras does) gives the same bits as the fused launch
        l1, g1 = fns[si].value_and_grad(yts[si], yps[si])
        assert l1.item() == loss[si]
        assert torch.equal(g1, dpreds[si])


def test_v1_vs_oracle():
    rng = np.random.default_rng(17)
    S, B, C, n = 7, 2, 20, 6
    yt = synth.make_labels(rng, n, [S], C, synth.ANCHORS_V2)[0]
    yp = rng.uniform(0.02, 0.98, (n, S, S, B * 5 + C)).astype(np.float32)
    kw = dict(binary_weight=0.5, loss_weight=[5, 5, 1, 1])
    fn = wrap(1)((S, S), B, C, **kw)
    loss, grad = fn.value_and_grad(yt, yp)
    l_ref, g_ref, _ = ol.loss_and_grad(ol.GridLossSpec(version=1, grid_shape=(S, S), bbox_num=B, class_num=C, **kw), yt, yp)
    assert abs(float(loss) - l_ref) <= LOSS_RTOL * abs(l_ref)
    check_grad(grad, g_ref, "v1")


def test_variants_v4_v3():
    cfg = synth.make_config("v4-608", batch=2, seed=21)
    S, B, C = 38, 3, 80
    yt, yp = cfg["y_trues"][1], cfg["y_preds"][1]
    anc = cfg["anchors"][3:6]
    for ver, kw in [
        (4, dict(anchors=anc, loss_weight=[2, 3, 0.5], binary_weight=np.array([0.4]), wh_reg_weight=0.05,
                 ignore_thresh=0.5, truth_thresh=0.7, label_smooth=0.05, focal_loss_gamma=1.5)),
        (4, dict(anchors=None, loss_weight=[1, 1, 1], focal_loss_gamma=1)),
        (3, dict(anchors=anc, loss_weight=[1, 2, 3, 4], use_focal_loss=True, focal_loss_gamma=2,
                 use_scale=False, binary_weight=0.5, ignore_thresh=0.4)),
        (3, dict(anchors=anc, loss_weight=[1, 1, 1, 1], use_focal_loss=True, focal_loss_gamma=3)),
    ]:
        fn = wrap(ver)((S, S), B, C, **kw)
        loss, grad = fn.value_and_grad(yt, yp)
        l_ref, g_ref, _ = ol.loss_and_grad(ol.GridLossSpec(version=ver, grid_shape=(S, S), bbox_num=B, class_num=C, **kw), yt, yp)
        assert abs(float(loss.reshape(-1)[0]) - l_ref) <= LOSS_RTOL * abs(l_ref), kw
        check_grad(grad, g_ref, str(kw))


def test_ragged_and_unaligned_inputs():
    """n_cells not a multiple of the tile / of 4, and base pointers off 16-byte alignment."""
    cfg = synth.make_config("v4-608", batch=1, seed=5)   # 361 cells: ragged tail
    S, B, C = 19, 3, 80
    yt, yp = cfg["y_trues"][0], cfg["y_preds"][0]
    fn = wrap(4)((S, S), B, C, anchors=cfg["anchors"][:3], loss_weight=[1, 5, 1])
    l_ref, g_ref, _ = ol.loss_and_grad(ol.GridLossSpec(version=4, grid_shape=(S, S), bbox_num=B, class_num=C,
                                                       anchors=cfg["anchors"][:3], loss_weight=[1, 5, 1]), yt, yp)
    loss, grad = fn.value_and_grad(yt, yp)
    assert abs(float(loss) - l_ref) <= LOSS_RTOL * abs(l_ref)
    check_grad(grad, g_ref, "ragged")
    # unaligned: place the tensors 4 bytes into a larger buffer
    def shifted(a):
        buf = torch.empty(a.size + 1, dtype=torch.float32, device="cuda")
        v = buf[1:].view(*a.shape)
        v.copy_(torch.from_numpy(a))
        assert v.data_ptr() % 16 != 0
        return v
    syt, syp = shifted(yt), shifted(yp)
    l2, d2, _ = engine.loss_fwd_bwd([fn.params], [syt], [syp])
    assert l2.item() == float(loss)
    assert np.array_equal(d2[0].cpu().numpy(), grad)


def test_forward_only_and_autograd_and_global_batch():
    cfg = synth.make_config("v3-416", batch=2, seed=8)
    S, B, C = 13, 3, 80
    fn = wrap(3)((S, S), B, C, anchors=cfg["anchors"][:3])
    yt = torch.from_numpy(cfg["y_trues"][0]).cuda()
    yp = torch.from_numpy(cfg["y_preds"][0]).cuda()
    l_fwd, none, _ = engine.loss_fwd_bwd([fn.params], [yt], [yp], want_grad=False)
    assert none is None
    l, g = fn.value_and_grad(yt, yp)
    assert l_fwd.item() == l.item()
    # autograd contract: backward returns upstream * dpred
    ypg = yp.clone().requires_grad_()
    out = fn(yt, ypg)
    (out * 3.0).backward()
    assert torch.allclose(ypg.grad, 3.0 * g, rtol=1e-6, atol=0)
    # numpy in -> numpy scalar out
    out_np = fn(cfg["y_trues"][0], cfg["y_preds"][0])
    assert isinstance(out_np, np.ndarray) and out_np.shape == () and out_np == l.item()
    # sharded batch: each half with global_batch=2 sums to the full-batch loss, grads identical
    halves = []
    for i in range(2):
        li, di, _ = engine.loss_fwd_bwd([fn.params], [yt[i:i + 1].contiguous()], [yp[i:i + 1].contiguous()],
                                        global_batch=2)
        halves.append((li.item(), di[0]))
    assert abs(halves[0][0] + halves[1][0] - l.item()) <= 2e-6 * abs(l.item())
    assert torch.equal(torch.cat([halves[0][1], halves[1][1]]), g)


def test_error_behaviour():
    fn = wrap(4)((19, 19), 3, 80)
    with pytest.raises(ValueError):
        fn(torch.zeros(1, 19, 19, 85, device="cuda"), torch.zeros(1, 19, 19, 254, device="cuda"))
    with pytest.raises(Exception):
        engine.loss_fwd_bwd([fn.params], [torch.zeros(1, 19, 19, 85)], [torch.zeros(1, 19, 19, 255)])


def test_grid_iou_entry():
    cfg = synth.make_config("v4-608", batch=2, seed=3)
    S, B, C = 19, 3, 80
    yt = torch.from_numpy(cfg["y_trues"][0]).cuda().reshape(2, S, S, 1, 85)
    yp = torch.from_numpy(cfg["y_preds"][0]).cuda().reshape(2, S, S, B, 85)
    from tf2_yolo_b200.yolov4.losses import cal_iou
    iou, ciou = cal_iou(yt[..., :4], yp[..., :4], (S, S), return_ciou=True)
    ri, rc = ol.grid_iou(torch.from_numpy(cfg["y_trues"][0]).double().reshape(2, S, S, 1, 85)[..., :4],
                         torch.from_numpy(cfg["y_preds"][0]).double().reshape(2, S, S, B, 85)[..., :4],
                         (S, S), want_ciou=True)
    assert np.allclose(iou.cpu().numpy(), ri.numpy(), rtol=1e-5, atol=1e-6)
    obj = cfg["y_trues"][0][..., 4] == 1      # CIoU is only defined where a label box exists
    assert np.allclose(ciou.cpu().numpy()[obj], rc.numpy()[obj], rtol=1e-5, atol=1e-6)


def test_fused_loss_decode_equals_separate_calls():
    """yb_loss_decode_fused: y_pred read once; loss, gradient and decoded rows identical to the two calls."""
    for name, ver, thr in (("v4-608", 4, 0.5), ("v3-416", 3, 0.3), ("v2-416", 2, 0.4)):
        cfg = synth.make_config(name, batch=5, seed=61)
        B, C = cfg["bbox_num"], cfg["class_num"]
        fns = []
        for si, S in enumerate(cfg["grids"]):
            kw = dict(anchors=cfg["anchors"][si * B:(si + 1) * B])
            kw["loss_weight"] = [1, 5, 1] if ver == 4 else [1, 1, 5, 1]
            fns.append(wrap(ver)((S, S), B, C, **kw))
        yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
        yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
        loss0, d0, _ = fused_losses(fns, yts, yps)
        rows0, off0 = engine.decode_batch(yps, C, thr, ver, capacity=5 * 4096)
        loss1, d1, _, rows1, off1 = engine.loss_decode_fused([f.params for f in fns], yts, yps, thr, capacity=5 * 4096)
        assert torch.equal(loss0, loss1)
        assert all(torch.equal(a, b) for a, b in zip(d0, d1))
        assert torch.equal(off0, off1)
        n = int(off0[-1])
        assert n > 0 and torch.equal(rows0[:n], rows1[:n])


def test_from_logits_head_fusion():
    """Raw head outputs in, d/d(raw) out: equals the oracle loss composed with sigmoid / anchor*exp."""
    from tf2_yolo_b200.grid_loss import wrap_yolo_loss_from_logits
    for ver, name in ((4, "v4-608"), (3, "v3-416")):
        cfg = synth.make_config(name, batch=2, seed=71)
        S, B, C = cfg["grids"][1], cfg["bbox_num"], cfg["class_num"]
        anc = cfg["anchors"][B:2 * B]
        act = cfg["y_preds"][1].astype(np.float64).reshape(-1, B, 5 + C)
        raw = act.copy()                                   # invert the head transform of the synthetic outputs
        raw[..., 0:2] = np.log(act[..., 0:2] / (1 - act[..., 0:2]))
        raw[..., 4:] = np.log(act[..., 4:] / (1 - act[..., 4:]))
        raw[..., 2:4] = np.log(act[..., 2:4] / anc[None])
        raw = raw.reshape(cfg["y_preds"][1].shape).astype(np.float32)
        kw = dict(anchors=anc, loss_weight=[1, 5, 1] if ver == 4 else [1, 1, 5, 1])
        fn = wrap_yolo_loss_from_logits(ver, (S, S), B, C, **kw)
        loss, grad = fn.value_and_grad(cfg["y_trues"][1], raw)
        spec = ol.GridLossSpec(version=ver, grid_shape=(S, S), bbox_num=B, class_num=C, **kw)
        l_ref, g_ref = ol.loss_and_grad_from_logits(spec, cfg["y_trues"][1], raw)
        assert abs(float(loss) - l_ref) <= LOSS_RTOL * abs(l_ref), (ver, float(loss), l_ref)
        check_grad(grad, g_ref, f"logits v{ver}")
    with pytest.raises(ValueError):
        wrap_yolo_loss_from_logits(2, (13, 13), 5, 20, anchors=synth.ANCHORS_V2)


def test_loss_bits_do_not_depend_on_the_schedule():
    """Tiles are handed out dynamically, so which CTA sums which cells changes from launch to
    launch; the binned (exact) accumulation must still give identical bits every time, for the
    loss, every term and the in-training metric sums."""
    cfg = synth.make_config("v4-608", batch=16, seed=11)
    B, C = cfg["bbox_num"], cfg["class_num"]
    fns = [wrap(4)((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    yts = [torch.from_numpy(a).cuda() for a in cfg["y_trues"]]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    ref = None
    for _ in range(6):
        loss, dpreds, terms, metrics = fused_losses(fns, yts, yps, want_terms=True, want_metrics=True)
        cur = (loss.clone(), terms.clone(), metrics.clone(), [d.clone() for d in dpreds])
        if ref is None:
            ref = cur
            continue
        assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and torch.equal(cur[2], ref[2])
        assert all(torch.equal(a, b) for a, b in zip(cur[3], ref[3]))
    # the terms agree with an fp64 sum of the same addends far below the loss tolerance
    spec = ol.GridLossSpec(version=4, grid_shape=(cfg["grids"][2],) * 2, bbox_num=B, class_num=C,
                           anchors=cfg["anchors"][2 * B:3 * B], loss_weight=[1, 5, 1])
    l_ref, _, _ = ol.loss_and_grad(spec, cfg["y_trues"][2], cfg["y_preds"][2])
    assert abs(ref[1][2, 0].item() - l_ref) <= LOSS_RTOL * abs(l_ref)


def test_nonfinite_inputs_propagate():
    """An Inf / NaN head output must poison the loss like it does in the reference (sums of
    non-finite addends), not vanish in the binned accumulation."""
    cfg = synth.make_config("v3-416", batch=2, seed=12)
    B, C = cfg["bbox_num"], cfg["class_num"]
    S = cfg["grids"][0]
    fn = wrap(3)((S, S), B, C, anchors=cfg["anchors"][:B])
    yt = torch.from_numpy(cfg["y_trues"][0]).cuda()
    yp = torch.from_numpy(cfg["y_preds"][0]).cuda().clone()
    yp[0, 0, 0, 2] = float("nan")     # width of box 0 in cell (0, 0)
    loss, _ = fn.value_and_grad(yt, yp)
    assert not np.isfinite(float(loss.reshape(-1)[0]))
