"""Device-side ingest (the first "next" row of the scope table): BGR->RGB + cv2-exact INTER_LINEAR resize of whole frames, and
the fused pipeline on frames that do not have the network size, against cv2 / the oracle."""
import numpy as np
import pytest
import torch

from util import DEV
from yolo_deepsort_b200._lib import check, lib, ptr, stream_ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1080, 1920, 608, 608), (720, 1280, 608, 608), (1216, 1216, 608, 608), (480, 640, 416, 416),
                                   (300, 500, 608, 608), (608, 608, 416, 416), (97, 131, 64, 48), (416, 416, 416, 416)])
@pytest.mark.parametrize("swap", [0, 1])
def test_resize_bit_exact_vs_cv2(shape, swap):
    import cv2
    from oracle.cv_resize_ref import resize_linear_u8
    h, w, H, W = shape
    rng = np.random.default_rng(h * 7 + w + swap)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dst = torch.zeros((H, W, 3), dtype=torch.uint8, device=DEV)
    check(lib().ydst_resize_u8(ptr(torch.from_numpy(src).to(DEV)), h, w, ptr(dst), H, W, swap, stream_ptr()))
    ref = cv2.resize(src[:, :, ::-1] if swap else src, (W, H), interpolation=cv2.INTER_LINEAR)
    np.testing.assert_array_equal(resize_linear_u8(np.ascontiguousarray(src[:, :, ::-1]) if swap else src, W, H), ref)   # the oracle is pinned by cv2
    np.testing.assert_array_equal(dst.cpu().numpy(), ref)


def test_pipeline_on_frames_of_another_size():
    """A 500x700 BGR clip through the fused pipeline vs the reference flow restated with cv2 + the oracle: resize to the
    network size, detect, scale the boxes back (resize_boxes), crop the ORIGINAL frame for ReID, track."""
    import cv2
    from oracle import darknet_ref as D, reid_ref as R, sort_ref as S
    from oracle.synth import make_frame
    from test_gpu_pipeline import build
    big = [make_frame(500, 700, seed=s) for s in (0, 1)]
    small = [cv2.resize(f, (416, 416), interpolation=cv2.INTER_LINEAR) for f in big]
    model, blocks, ws, sd, ds, pipe = build(small)
    clip = [big[0]] * 5 + [big[1]] * 1
    orc = S.DeepSortRef(lambda fr, tl: R.extract(sd, fr, tl), max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
    res = []
    for f in clip:
        pipe.submit(np.ascontiguousarray(f[:, :, ::-1]), bgr=True)      # BGR, as cv2.VideoCapture delivers it
        res.append(pipe.collect())
    n_rows = 0
    for t, (f, (tracks, dets)) in enumerate(zip(clip, res)):
        image = cv2.resize(f, (416, 416), interpolation=cv2.INTER_LINEAR)
        det = D.detect(blocks, ws, image, (416, 416), 0.5, 0.4)
        assert det is not None and dets.shape == det.shape, f"frame {t}"
        det[:, 0] *= np.float32(700 / 416); det[:, 2] *= np.float32(700 / 416)
        det[:, 1] *= np.float32(500 / 416); det[:, 3] *= np.float32(500 / 416)
        # same detections in the same order (the calibrated heads keep the score order clear of rounding noise); the boxes are
        # scaled by 700/416 and 500/416 here, so their corners leave the half-pixel lattice and a crop may move by one pixel
        size = np.minimum(det[:, 2] - det[:, 0], det[:, 3] - det[:, 1])
        rel = np.abs(dets[:, :4] - det[:, :4]).max(1) / size
        assert rel.max() < 1e-2 and (dets[:, 5] == det[:, 5]).all(), f"frame {t}: boxes {rel.max():.3g}"
        tlwh, conf, cls = D.to_tracker_inputs(det, [0, 2, 4])
        ref = np.asarray(orc.update(tlwh, conf, f, torch.from_numpy(cls)), np.int32).reshape(-1, 6)
        got = np.asarray(tracks, np.int32).reshape(-1, 6)
        assert got.shape == ref.shape, f"frame {t}: {got.shape} vs {ref.shape}"
        if t < 5:                                           # one held scene: ids and classes identical, in order
            np.testing.assert_array_equal(got[:, 4:], ref[:, 4:], err_msg=f"frame {t}")
            assert np.abs(got[:, :4] - ref[:, :4]).max(initial=0) <= 2, f"frame {t}"
        n_rows += len(ref)
    assert n_rows > 50
