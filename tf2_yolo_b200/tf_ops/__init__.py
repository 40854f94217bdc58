"""TensorFlow custom-op boundary (source + bindings; see yolo_loss_op.cc / yolo_loss_op.py)."""
