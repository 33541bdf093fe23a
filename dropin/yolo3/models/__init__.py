"""yolo3.models.Darknet (yolo3/models/models.py:277-366) -> yolo_deepsort_b200.Darknet (libydst, sm_100a)."""
from yolo_deepsort_b200.darknet import Darknet  # noqa: F401
