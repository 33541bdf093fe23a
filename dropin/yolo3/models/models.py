from yolo_deepsort_b200.darknet import Darknet, parse_model_config  # noqa: F401
