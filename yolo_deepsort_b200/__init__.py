"""B200-native (sm_100a) detect-and-track hot path behind the GlassyWing/yolo_deepsort Python surface.

    from yolo_deepsort_b200 import Darknet, DeepSort, VideoDetector

Every compute stage lives in libydst.so (hand-written CUDA, C ABI in include/ydst.h); this package is the host-side
mirror of the reference interface.  There is no CPU fallback.
"""
from .darknet import Darknet, parse_model_config, soft_non_max_suppression, resize_boxes, p1p2Toxywh  # noqa: F401
from .reid import Extractor  # noqa: F401
from .deepsort import DeepSort  # noqa: F401
from .pipeline import FramePipeline  # noqa: F401
from .detect import ImageDetector, VideoDetector  # noqa: F401
from .action import ActionIdentify  # noqa: F401
