from yolo_deepsort_b200.detect import ImageDetector  # noqa: F401
