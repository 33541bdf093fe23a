"""yolo3.detect.img_detect.ImageDetector (yolo3/detect/img_detect.py:37-153)."""
from yolo_deepsort_b200.detect import ImageDetector  # noqa: F401
