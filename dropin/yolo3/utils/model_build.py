"""yolo3.utils.model_build hot-path functions (yolo3/utils/model_build.py:12-19,52-137,326-332)."""
from yolo_deepsort_b200.darknet import soft_non_max_suppression, resize_boxes, p1p2Toxywh  # noqa: F401
