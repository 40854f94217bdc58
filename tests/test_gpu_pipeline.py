"""GPU: the pipelined host-buffer step (HostBatchStep) - box lists + pinned head outputs in, loss,
gradient (device) and NMS survivors (host) out - equals the separate calls on device tensors,
for any chunking, and falls back to the general chain when an image exceeds rows_per_img."""
import numpy as np
import pytest
import torch

from oracle import tools as ot
from tf2_yolo_b200 import engine, synth
from tf2_yolo_b200.grid_loss import fused_losses
from tf2_yolo_b200.pipeline import HostBatchStep
from tf2_yolo_b200.yolov4.losses import wrap_yolo_loss

pytestmark = pytest.mark.gpu


def setup(batch, seed):
    cfg = synth.make_config("v4-608", batch=batch, seed=seed)
    B, C = 3, 80
    fns = [wrap_yolo_loss((S, S), B, C, anchors=cfg["anchors"][si * B:(si + 1) * B], loss_weight=[1, 5, 1])
           for si, S in enumerate(cfg["grids"])]
    boxes, offs = synth.boxes_from_labels(cfg["y_trues"][-1], (608, 608))
    host_p = [torch.from_numpy(a).pin_memory() for a in cfg["y_preds"]]
    return cfg, fns, boxes, offs, host_p


def reference(cfg, fns, boxes, offs):
    labels = ot.encode_label_pyramid(boxes, offs, (608, 608), (76, 76), 80, 3, dtype=np.float32)
    yts = [torch.from_numpy(a).cuda() for a in labels]
    yps = [torch.from_numpy(a).cuda() for a in cfg["y_preds"]]
    loss, dpreds, _ = fused_losses(fns, yts, yps)
    rows, ro = engine.decode_batch_exact(yps, 80, 0.5, 4)
    g = engine.nms_batch(rows, ro, 80, 0.45, 2)
    n = int(g["out_offsets"][-1].item())
    return loss.cpu().numpy(), dpreds, g["out_rows"][:n].cpu().numpy(), g["out_offsets"].cpu().numpy()


@pytest.mark.parametrize("batch,chunks", [(12, 3), (12, 1), (7, 4), (5, 8)])
def test_pipelined_step_equals_the_separate_calls(batch, chunks):
    cfg, fns, boxes, offs, host_p = setup(batch, 31)
    loss_ref, d_ref, rows_ref, offs_ref = reference(cfg, fns, boxes, offs)
    pipe = HostBatchStep(fns, (608, 608), batch, 0.5, 0.45, 2, n_chunks=chunks, rows_per_img=1024,
                         max_boxes_per_img=int(np.diff(offs).max()), max_boxes=len(boxes))
    hb, ho = torch.from_numpy(boxes).pin_memory(), torch.from_numpy(offs).pin_memory()
    for _ in range(2):                                   # twice: staging buffers and control blocks are reused
        res = pipe.run(host_p, hb, ho)
        assert np.allclose(res["loss"], loss_ref, rtol=2e-6)
        if chunks == 1:
            assert np.array_equal(res["loss"], loss_ref)
        for a, b in zip(res["dpred"], d_ref):
            assert torch.equal(a, b)
        got = np.concatenate([r for r, _ in res["rows"]], axis=0)
        assert res["n_rows"] == len(rows_ref) and np.array_equal(got, rows_ref)
        per_img = np.concatenate([np.diff(o) for _, o in res["rows"]])
        assert np.array_equal(per_img, np.diff(offs_ref))


def test_chunk_with_a_crowded_image_is_redone_by_the_general_chain():
    cfg, fns, boxes, offs, host_p = setup(8, 32)
    loss_ref, d_ref, rows_ref, offs_ref = reference(cfg, fns, boxes, offs)
    per_img = np.diff(engine.decode_batch_exact([torch.from_numpy(a).cuda() for a in cfg["y_preds"]], 80, 0.5, 4)[1].cpu().numpy())
    cap = int(max(32, np.sort(per_img)[-2]))            # the most crowded image does not fit
    assert per_img.max() > cap
    pipe = HostBatchStep(fns, (608, 608), 8, 0.5, 0.45, 2, n_chunks=2, rows_per_img=cap,
                         max_boxes_per_img=int(np.diff(offs).max()), max_boxes=len(boxes))
    res = pipe.run(host_p, torch.from_numpy(boxes).pin_memory(), torch.from_numpy(offs).pin_memory())
    assert int(pipe.overflow_host.sum()) >= 1
    got = np.concatenate([r for r, _ in res["rows"]], axis=0)
    assert np.array_equal(got, rows_ref)
    assert np.allclose(res["loss"], loss_ref, rtol=2e-6)


def test_bad_boxes_raise_like_the_reference():
    cfg, fns, boxes, offs, host_p = setup(4, 33)
    bad = boxes.copy()
    bad[0, 4] = 80                                       # class index out of range: IndexError in the reference
    pipe = HostBatchStep(fns, (608, 608), 4, n_chunks=2, max_boxes_per_img=int(np.diff(offs).max()), max_boxes=len(boxes))
    with pytest.raises(IndexError):
        pipe.run(host_p, torch.from_numpy(bad).pin_memory(), torch.from_numpy(offs).pin_memory())
