"""Raw-logit replacement for the activation tail of the reference's YOLOv3 / YOLOv4 heads
(SURVEY.md 8f row 2).

The reference head (yolov4/models/__init__.py:37-66, yolov3/models/__init__.py:36-63) ends every
box in four 1x1 convolutions followed by activations - sigmoid (xy, objectness, class scores),
``exp`` times the anchor (wh: the ``Anchor`` layer of yolov4/models/backbone.py:40-60, or
``activation='exponential'`` + ``Multiply`` in v3) - and a per-scale ``Concatenate``.  The loss
then reads the activated tensor and autodiff walks back through the activations.

``yolo_head_logits`` builds the SAME convolutions under the SAME layer names (so the weights of a
model built by the reference load by name) but without the activations and the anchor layers:
the model emits raw ``[tx, ty, tw, th, tc, tp_0 .. tp_{C-1}]`` per box, which is exactly the
layout ``from_logits`` of the fused loss kernel takes (yb_loss_params.from_logits): the kernel
applies the head transform itself and returns dL/d(raw), saving one write + read of every head
tensor and the activation backward passes.  ``activate_head`` is the inference-side transform for
callers that want the reference's activated tensors (e.g. to feed ``utils.tools.decode``).

TensorFlow is imported lazily: this module is importable (and tested with a stand-in) without it.
"""
import numpy as np


def yolo_head_logits(model_body, class_num=80, anchors=None, conv_layer=None):
    """Keras model with the reference head's convolutions and names, emitting raw logits.

    ``model_body``: the backbone model whose outputs feed the head (as in ``yolo_head``).
    ``anchors`` only decides how many boxes each output tensor carries (len(anchors) / outputs), as
    in the reference; their values are not used here - they go to the loss (``anchors=`` of
    ``wrap_yolo_loss_from_logits``) and to ``activate_head``.
    ``conv_layer``: the reference's ``DarknetConv2D`` factory (defaults to importing it from the
    reference's ``yolov4.models.backbone``)."""
    from tensorflow.keras.layers import Concatenate
    from tensorflow.keras.models import Model
    if conv_layer is None:
        from yolov4.models.backbone import DarknetConv2D as conv_layer
    if anchors is None:
        raise ValueError("anchors decide the number of boxes per output tensor")
    anchors = np.array(anchors)
    out_tensors = model_body.output
    tensor_num = len(out_tensors)
    if len(anchors) % tensor_num > 0:
        raise ValueError(
            "The total number of anchor boxs "
            "should be a multiple of the number "
            f"{tensor_num} of output tensors")
    abox_num = len(anchors)//tensor_num
    outputs_list = []
    for i_tensor, out_tensor in enumerate(out_tensors):
        output_list = []
        for i_box in range(abox_num):
            prefix = f"out{i_tensor + 1}_box{i_box + 1}"
            output_list += [conv_layer(2, 1, name=f"{prefix}_xy_conv")(out_tensor),        # no sigmoid
                            conv_layer(2, 1, name=f"{prefix}_wh_conv")(out_tensor),        # no exp, no anchor
                            conv_layer(1, 1, name=f"{prefix}_conf_conv")(out_tensor),      # no sigmoid
                            conv_layer(class_num, 1, name=f"{prefix}_prob_conv")(out_tensor)]
        outputs_list.append(Concatenate(name=f"out{i_tensor + 1}_concat")(output_list))
    return Model(model_body.input, outputs_list)


def activate_head(raw, anchors, class_num):
    """The reference head's activations on one raw output tensor (N, S, S, B*(5+C)):
    sigmoid xy / objectness / class scores, ``anchor * exp`` for wh -> the activated tensor the
    reference's model emits (same layout).  ``anchors``: the (B, 2) anchors of this scale."""
    import tensorflow as tf
    anchors = tf.reshape(tf.constant(np.asarray(anchors, dtype=np.float32)), (1, 1, 1, -1, 2))
    shape = tf.shape(raw)
    cell = tf.reshape(raw, (shape[0], shape[1], shape[2], -1, 5 + class_num))
    out = tf.concat([tf.sigmoid(cell[..., 0:2]), tf.exp(cell[..., 2:4]) * anchors, tf.sigmoid(cell[..., 4:])], axis=-1)
    return tf.reshape(out, shape)


def wrap_yolo_loss_from_logits(version, grid_shape, bbox_num, class_num, anchors, **kwargs):
    """``yolo_loss(y_true, raw)`` for the outputs of ``yolo_head_logits``: the v3 / v4 loss with the
    head transform inside the kernel (the YoloGridLoss op with ``from_logits=True``).  Keyword
    arguments as the reference's ``wrap_yolo_loss`` of that version."""
    from .yolo_loss_op import make_loss
    if version not in (3, 4):
        raise ValueError("from-logits heads exist for yolov3 / yolov4 (v1 / v2 end in a softmax)")
    if anchors is None:
        raise ValueError("from-logits needs the anchors of the scale")
    return make_loss(version, grid_shape, bbox_num, class_num, anchors=anchors, from_logits=True, **kwargs)
