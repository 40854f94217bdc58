"""Drop-in for the hot-path functions of the reference's ``utils`` package:
``tools.decode / nms / cal_iou``, ``kmeans.kmeans`` and ``measurement.PRfunc /
create_score_mat`` -- same names and signatures, computed by CUDA kernels."""
from . import kmeans, measurement, tools  # noqa: F401
