"""Drop-in for the hot path of the reference's ``yolov3`` package (losses only)."""
