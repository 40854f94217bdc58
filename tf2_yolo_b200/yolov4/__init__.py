"""Drop-in for the hot path of the reference's ``yolov4`` package (losses only)."""
