"""yolo3.detect.video_detect.VideoDetector (yolo3/detect/video_detect.py:39-208)."""
from yolo_deepsort_b200.detect import VideoDetector  # noqa: F401
