"""Drop-in for the hot path of the reference's ``yolov1_5`` package (losses only)."""
