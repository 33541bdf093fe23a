"""deep_sort.DeepSort / build_tracker (deep_sort/__init__.py:1-11, deep_sort/deep_sort.py:16-88)."""
from yolo_deepsort_b200.deepsort import DeepSort  # noqa: F401

__all__ = ['DeepSort', 'build_tracker']


def build_tracker(cfg, use_cuda):
    d = cfg.DEEPSORT
    return DeepSort(d.REID_CKPT, max_dist=d.MAX_DIST, min_confidence=d.MIN_CONFIDENCE, nms_max_overlap=d.NMS_MAX_OVERLAP,
                    max_iou_distance=d.MAX_IOU_DISTANCE, max_age=d.MAX_AGE, n_init=d.N_INIT, nn_budget=d.NN_BUDGET,
                    use_cuda=use_cuda)
