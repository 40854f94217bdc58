"""Host-side mirror of the reference's ``wrap_yolo_loss`` closures.

``GridLoss`` objects are what ``wrap_yolo_loss(...)`` returns in this package:
callables ``yolo_loss(y_true, y_pred)`` with the reference's meaning
(yolov4/losses/loss.py:75-167 and the v1/v2/v3 siblings), computed by the fused
CUDA kernel behind ``yb_loss_fwd_bwd``.  The gradient w.r.t. ``y_pred`` is
produced in the same pass and handed to the caller's autodiff:

* torch CUDA tensors (the carrier used in this repo): a ``torch.autograd.Function``
  whose backward returns ``upstream * dpred`` -- the same contract as the
  ``@tf.RegisterGradient`` of the TensorFlow custom op (tf_ops/).
* host arrays (numpy / anything ``np.asarray`` accepts): copied to the GPU,
  evaluated, and the scalar copied back (``value_and_grad`` also returns the
  gradient as a numpy array).

There is no CPU implementation here: without a CUDA device every call raises.
"""
import numpy as np
import torch

from . import engine


class _GridLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred, spec, global_batch):
        need_grad = ctx.needs_input_grad[1]
        loss, dpreds, _ = engine.loss_fwd_bwd([spec.params], [y_true], [y_pred],
                                              global_batch=global_batch, want_grad=need_grad)
        if need_grad:
            ctx.save_for_backward(dpreds[0])
        return loss.reshape(spec.out_shape)

    @staticmethod
    def backward(ctx, grad_out):
        (dpred,) = ctx.saved_tensors
        return None, dpred * grad_out.reshape(-1)[0], None, None


class GridLoss:
    """Callable returned by ``wrap_yolo_loss``; keeps the keyword surface as attributes."""

    def __init__(self, version, grid_shape, bbox_num, class_num, **kw):
        self.version = version
        self.grid_shape = tuple(int(g) for g in grid_shape)
        self.bbox_num = int(bbox_num)
        self.class_num = int(class_num)
        self.kwargs = dict(kw)
        bw = kw.get("binary_weight", 1)
        # a 1-element ndarray binary_weight makes the reference return shape (1,)
        self.out_shape = tuple(np.shape(bw)) if isinstance(bw, np.ndarray) else ()
        self.params = engine.make_loss_params(version, grid_shape, bbox_num, class_num, **kw)
        self.global_batch = None   # set to the global batch when the batch is sharded over ranks
        self.__name__ = "yolo_loss"

    # -- helpers ---------------------------------------------------------------
    def _to_device(self, a, device):
        if torch.is_tensor(a):
            t = a
        elif hasattr(a, "__dlpack__") and not isinstance(a, np.ndarray):
            t = torch.from_dlpack(a)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
        if not t.is_cuda:
            t = t.to(device, non_blocking=True)
        if t.dtype != torch.float32:
            t = t.float()          # Keras casts y_true to y_pred.dtype; heads are fp32
        return t.contiguous()

    def _device(self, *arrays):
        for a in arrays:
            if torch.is_tensor(a) and a.is_cuda:
                return a.device
        if not torch.cuda.is_available():
            raise engine.N.YoloB200Error("no CUDA device: tf2_yolo_b200 has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    # -- the reference's call signature ---------------------------------------
    def __call__(self, y_true, y_pred):
        host_in = not (torch.is_tensor(y_pred) and y_pred.is_cuda)
        dev = self._device(y_true, y_pred)
        yt = self._to_device(y_true, dev)
        yp = self._to_device(y_pred, dev)
        if torch.is_tensor(y_pred) and y_pred.requires_grad and yp is not y_pred and not yp.requires_grad:
            yp = yp.requires_grad_()
        out = _GridLossFn.apply(yt, yp, self, self.global_batch)
        if host_in and not torch.is_tensor(y_pred):
            return out.detach().cpu().numpy()
        return out

    def value_and_grad(self, y_true, y_pred):
        """(loss, dL/dy_pred) in the caller's container type."""
        host_in = not torch.is_tensor(y_pred)
        dev = self._device(y_true, y_pred)
        yt = self._to_device(y_true, dev)
        yp = self._to_device(y_pred, dev)
        loss, dpreds, _ = engine.loss_fwd_bwd([self.params], [yt], [yp], global_batch=self.global_batch)
        loss = loss.reshape(self.out_shape)
        grad = dpreds[0].reshape(tuple(np.shape(y_pred)) if host_in else y_pred.shape)
        if host_in:
            return loss.cpu().numpy(), grad.cpu().numpy()
        return loss, grad


def fused_losses(loss_fns, y_trues, y_preds, global_batch=None, want_grad=True, want_terms=False,
                 dpreds=None, want_metrics=False, recall_iou_threshold=0.5):
    """All FPN scales of one train step in ONE launch (what Keras does with
    ``loss=[f0, f1, f2]`` in three).  Tensors must already be fp32 CUDA.
    Returns (loss per scale [n] fp32 CUDA, dpreds, terms[, metrics]); with ``want_metrics`` the
    in-training metrics (obj_acc, mean_iou, class_acc, recall + raw sums, per scale) come out of
    the same pass."""
    return engine.loss_fwd_bwd([f.params for f in loss_fns], list(y_trues), list(y_preds),
                               global_batch=global_batch, want_grad=want_grad,
                               want_terms=want_terms, dpreds=dpreds, want_metrics=want_metrics,
                               recall_iou_threshold=recall_iou_threshold)


def cal_iou_grid(xywh_true, xywh_pred, grid_shape, return_ciou=False):
    """``cal_iou`` of the loss modules (yolov4/losses/loss.py:10-61) on CUDA tensors."""
    if not (torch.is_tensor(xywh_pred) and xywh_pred.is_cuda):
        dev = torch.device("cuda", torch.cuda.current_device())
        xywh_true = torch.as_tensor(np.asarray(xywh_true), dtype=torch.float32).to(dev)
        xywh_pred = torch.as_tensor(np.asarray(xywh_pred), dtype=torch.float32).to(dev)
    bt = xywh_true.float().contiguous()
    bp = xywh_pred.float().contiguous()
    return engine.grid_iou(bt, bp, grid_shape, want_ciou=return_ciou)


def wrap_yolo_loss_from_logits(version, grid_shape, bbox_num, class_num, **kwargs):
    """Extension: the same loss with the head transform folded in.

    ``yolo_loss(y_true, raw)`` takes the RAW outputs of the head's 1x1 convolutions
    (yolov4/models/__init__.py:42-60 without the activations / Anchor layer / per-box Concatenate
    re-ordering: layout [tx, ty, tw, th, tc, tp_0..tp_{C-1}] per box), applies sigmoid to
    xy / objectness / class scores and ``anchor * exp`` to wh inside the kernel, and returns the
    gradient with respect to the raw values - one read and one write of the head tensor instead
    of the activation round trip.  v3 / v4 only; ``anchors`` are required; the anchors are the
    constants the loss sees (not trainable through this entry point)."""
    if kwargs.get("anchors") is None:
        raise ValueError("from-logits needs the anchors of the scale")
    return GridLoss(version, grid_shape, bbox_num, class_num, from_logits=True, **kwargs)
