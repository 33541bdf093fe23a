"""Oracle: cv2.resize(uint8, INTER_LINEAR) restated in numpy integer arithmetic, plus the
DeepSort crop rule.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party algorithm: OpenCV 4.13.0 ``cv::resize`` (``resizeGeneric_`` with
``HResizeLinear`` / ``VResizeLinearVec_32s8u``, INTER_RESIZE_COEF_BITS = 11), called by the
reference at deep_sort/deep/feature_extractor.py:45 (crop -> 64x128) and
yolo3/detect/img_detect.py:70 (frame -> model size).  Not under /root/reference, so restated
from its published source semantics and pinned by live differential tests against the installed
cv2 (tests/test_oracle_thirdparty.py: 0 mismatching pixels on 400 random crop sizes).

Per axis (dst n, src s):  f = float32((d + 0.5) * (s / n) - 0.5),  i0 = floor(f),  fr = f - i0.
  x axis: if i0 < 0 -> (i0, fr) = (0, 0); if i0 >= s-1 -> (s-1, 0); i1 = min(i0+1, s-1)
  y axis: fr is NOT zeroed; row indices are clamped: r0 = clip(i0, 0, s-1), r1 = clip(i0+1, 0, s-1)
  weights are rounded independently: w1 = rint(fr * 2048), w0 = rint((1 - fr) * 2048)   (half-even)
Horizontal pass (int32):  Hrow = src[i0] * a0 + src[i1] * a1
Vertical pass:            dst = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
"""
import numpy as np


def axis_coeffs(n_dst, n_src, clamp_frac):
    d = np.arange(n_dst)
    f = ((d + 0.5) * (n_src / n_dst) - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int32)
    fr = (f - i0).astype(np.float32)
    if clamp_frac:                        # x axis
        lo = i0 < 0
        i0[lo] = 0; fr[lo] = 0
        hi = i0 >= n_src - 1
        i0[hi] = n_src - 1; fr[hi] = 0
        i1 = np.minimum(i0 + 1, n_src - 1)
    else:                                 # y axis
        i1 = np.clip(i0 + 1, 0, n_src - 1)
        i0 = np.clip(i0, 0, n_src - 1)
    w1 = np.rint(fr * np.float32(2048)).astype(np.int32)
    w0 = np.rint((np.float32(1) - fr) * np.float32(2048)).astype(np.int32)
    return i0, i1, w0, w1


def resize_linear_u8(src, dst_w, dst_h):
    """src (h,w,c) uint8 -> (dst_h,dst_w,c) uint8, identical to cv2.resize(src,(dst_w,dst_h),INTER_LINEAR)."""
    sh, sw = src.shape[:2]
    if (sh, sw) == (dst_h, dst_w):
        return src.copy()
    x0, x1, a0, a1 = axis_coeffs(dst_w, sw, True)
    y0, y1, b0, b1 = axis_coeffs(dst_h, sh, False)
    s = src.astype(np.int32)
    H = s[:, x0] * a0[None, :, None] + s[:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (H[y0] >> 4)) >> 16) + ((b1[:, None, None] * (H[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def crop_box(tlwh, width, height):
    """DeepSort._s_tlwh_to_xyxy (deep_sort/deep_sort.py:116-122): Python int() truncation toward
    zero of fp32 values; note the -1 and that x+w / y+h are fp32 sums."""
    x, y, w, h = (np.float32(v) for v in tlwh)
    x1 = max(int(x), 0)
    x2 = min(int(np.float32(x + w)), width - 1)
    y1 = max(int(y), 0)
    y2 = min(int(np.float32(y + h)), height - 1)
    return x1, y1, x2, y2


def crops_to_batch(frame_rgb_u8, boxes_tlwh):
    """_get_features crops + Extractor._preprocess (deep_sort/deep_sort.py:133-141,
    deep_sort/deep/feature_extractor.py:34-51): returns (m,3,128,64) float32, normalised."""
    H, W = frame_rgb_u8.shape[:2]
    mean = np.array([0.485, 0.456, 0.406], np.float32)[None, :, None, None]
    std = np.array([0.229, 0.224, 0.225], np.float32)[None, :, None, None]
    ims = []
    for b in np.asarray(boxes_tlwh, np.float32):
        x1, y1, x2, y2 = crop_box(b, W, H)
        crop = frame_rgb_u8[y1:y2, x1:x2]
        if crop.shape[0] == 0 or crop.shape[1] == 0:
            raise ValueError("empty crop (cv2.resize raises in the reference)")
        ims.append(resize_linear_u8(crop, 64, 128).astype(np.float32).transpose(2, 0, 1))
    batch = np.stack(ims, 0) / np.float32(255.)
    return (batch - mean) / std
