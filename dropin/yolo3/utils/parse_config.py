"""yolo3.utils.parse_config.parse_model_config (yolo3/utils/parse_config.py:1-19)."""
from yolo_deepsort_b200.darknet import parse_model_config  # noqa: F401
