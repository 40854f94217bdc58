"""Drop-in for the hot path of the reference's ``yolov2`` package (losses only)."""
