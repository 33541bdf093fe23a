from yolo_deepsort_b200.deepsort import DeepSort  # noqa: F401
