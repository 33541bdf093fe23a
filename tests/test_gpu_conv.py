"""Parity of the tcgen05 implicit-GEMM convolution (and the direct first-layer kernel) against a plain PyTorch fp32
convolution of the same fp16-rounded operands.  Tolerance: the kernel accumulates in fp32 and rounds the result to fp16
once, so |y - ref| <= 2e-3 * max(1, |ref|) (fp16 has 11 significand bits -> 4.9e-4 relative; the margin covers the
accumulation-order difference); fp32-output (YOLO head) cases use 1e-3 relative to the output scale."""
import zlib

import numpy as np
import pytest
import torch

from util import DEV, conv2d_abi, conv2d_ref

pytestmark = pytest.mark.gpu

CASES = [
    # name, N, H, W, cin, cout, k, stride, bn, act, res_mode, out_f32
    ("1x1_s1_19", 1, 19, 19, 64, 32, 1, 1, True, 1, 0, False),
    ("3x3_s1_38", 1, 38, 38, 64, 128, 3, 1, True, 1, 0, False),
    ("3x3_s1_bk32", 1, 40, 24, 32, 64, 3, 1, True, 1, 0, False),
    ("3x3_s1_bk16", 1, 26, 26, 16, 32, 3, 1, True, 1, 0, False),
    ("3x3_s2_76", 1, 76, 76, 64, 128, 3, 2, True, 1, 0, False),
    ("3x3_s2_bk32", 1, 64, 64, 32, 64, 3, 2, True, 2, 0, False),
    ("3x3_s2_odd_tiles", 2, 38, 38, 128, 256, 3, 2, True, 1, 0, False),
    ("1x1_s2_batch3", 3, 64, 32, 64, 128, 1, 2, True, 0, 0, False),
    ("3x3_s2_small_images_batch10", 10, 16, 8, 256, 512, 3, 2, True, 3, 0, False),
    ("1x1_s2_small_images_batch7", 7, 16, 8, 256, 512, 1, 2, False, 0, 0, False),
    ("3x3_s2_tiny_images_batch33", 33, 8, 8, 64, 64, 3, 2, True, 1, 0, False),
    ("head_255_f32", 1, 19, 19, 1024, 255, 1, 1, False, 0, 0, True),
    ("res_after_act", 1, 38, 38, 128, 256, 3, 1, True, 1, 1, False),
    ("res_before_relu_batch5", 5, 32, 16, 128, 128, 3, 1, True, 3, 2, False),
    ("mish_1x1", 1, 76, 76, 128, 64, 1, 1, True, 2, 0, False),
    ("deep_k_wide_n", 1, 19, 19, 512, 1024, 3, 1, True, 1, 0, False),
    ("reid_8x4_batch7", 7, 8, 4, 512, 512, 3, 1, True, 3, 0, False),
    ("3x3_s1_76_two_boxes", 1, 76, 76, 128, 256, 3, 1, True, 1, 1, False),
    ("3x3_s1_152_wide_halo", 1, 152, 152, 64, 128, 3, 1, True, 1, 0, False),
    ("1x1_s1_768_concat_k", 1, 38, 38, 768, 256, 1, 1, True, 1, 0, False),
    ("reid_64x32_batch3", 3, 64, 32, 64, 64, 3, 1, True, 3, 2, False),
    ("persistent_reid_64x32_batch12", 12, 64, 32, 64, 64, 3, 1, True, 3, 2, False),
    ("persistent_1x1_152", 1, 152, 152, 128, 64, 1, 1, True, 1, 0, False),
    ("persistent_two_n_tiles_batch20", 20, 32, 16, 128, 256, 3, 1, True, 1, 1, False),
    ("persistent_f32_out", 2, 152, 152, 64, 48, 1, 1, False, 0, 0, True),
    ("persistent_narrow_32", 2, 152, 152, 64, 32, 1, 1, True, 1, 0, False),
    ("persistent_mish_res", 6, 76, 76, 64, 64, 3, 1, True, 2, 1, False),
    ("persistent_linear_bias", 4, 76, 76, 128, 128, 1, 1, False, 0, 0, False),
    ("persistent_relu_res_two_groups", 40, 32, 16, 128, 128, 3, 1, True, 3, 2, False),
    ("mpair_reid_64x32_batch24_odd_tiles", 24, 64, 32, 64, 64, 3, 1, True, 3, 2, False),
    ("mpair_1x1_304_batch2", 2, 304, 304, 64, 64, 1, 1, True, 1, 0, False),
    ("mpair_3x3_76_batch8_two_n_tiles", 8, 76, 76, 128, 256, 3, 1, True, 1, 1, False),
    ("mpair_persistent_reid_32x16_batch60_res", 60, 32, 16, 128, 128, 3, 1, True, 3, 2, False),
    ("mpair_persistent_1x1_152_batch3_odd", 3, 152, 152, 128, 64, 1, 1, True, 1, 0, False),
    ("cta2_3x3_76_batch2_res_odd_tiles", 2, 76, 76, 128, 256, 3, 1, True, 1, 1, False),
    ("cta2_1x1_38_batch8", 8, 38, 38, 512, 256, 1, 1, True, 1, 0, False),
    ("cta2_mish_3x3_38_batch4_res", 4, 38, 38, 256, 512, 3, 1, True, 2, 1, False),
    ("cta2_reid_16x8_batch40_res", 40, 16, 8, 256, 256, 3, 1, True, 3, 2, False),
    ("cta2_3x3_152_two_boxes_bn128", 2, 152, 152, 64, 128, 3, 1, True, 1, 0, False),
    ("cta2_persistent_3x3_76_batch8_res_odd", 8, 76, 76, 128, 256, 3, 1, True, 1, 1, False),
    ("cta2_persistent_mish_38_batch8_res", 8, 38, 38, 256, 512, 3, 1, True, 2, 1, False),
    ("cta2_persistent_reid_8x4_batch200_res", 200, 8, 4, 512, 512, 3, 1, True, 3, 2, False),
    ("cta2_persistent_1x1_76_batch8", 8, 76, 76, 256, 128, 1, 1, True, 1, 0, False),
    ("tap_persistent_s2_bk32_batch6", 6, 128, 128, 32, 64, 3, 2, True, 1, 0, False),
    ("tap_persistent_s1_bk32_res_batch2", 2, 104, 104, 32, 64, 3, 1, True, 1, 1, False),
    ("tap_persistent_s2_two_n_tiles_batch40", 40, 64, 32, 64, 256, 3, 2, True, 3, 0, False),
    ("tap_persistent_1x1_s2_small_images_batch300", 300, 16, 8, 256, 512, 1, 2, False, 0, 0, False),
    ("tap_persistent_mish_s2_batch3", 3, 152, 152, 64, 128, 3, 2, True, 2, 0, False),
    ("first_s1", 1, 64, 48, 3, 32, 3, 1, True, 1, 0, False),
    ("first_s2_64", 2, 32, 32, 3, 64, 3, 2, True, 3, 0, False),
    ("first_bias", 1, 16, 16, 3, 16, 3, 1, False, 0, 0, False),
]


@pytest.fixture(autouse=True)
def _opt_in_tilings(request, monkeypatch):
    """By default the persistent tile loop only takes weight-stationary layers with many tiles, and M-pair tiles are opt-in
    (YDST_PERSISTENT=2 / YDST_MPAIR / YDST_FORCE_MPAIR): the cases named after them widen the planner's choice so that those
    kernel instantiations stay covered on small shapes."""
    name = request.node.name
    if "persistent" in name or "mpair" in name:
        monkeypatch.setenv("YDST_PERSISTENT", "2")
    if "tap_persistent" in name:
        monkeypatch.setenv("YDST_TAP_PERSISTENT", "2")
    if "cta2" in name:
        monkeypatch.setenv("YDST_CTA2", "2")
    if "mpair" in name:
        monkeypatch.setenv("YDST_MPAIR", "1")
        monkeypatch.setenv("YDST_FORCE_MPAIR", "2")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_parity(case):
    name, N, H, W, cin, cout, k, stride, use_bn, act, res_mode, out_f32 = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    x = torch.randn(N, H, W, cin, generator=g)
    x = (x.to(DEV) if cin == 3 else x.half().to(DEV))
    w = (torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))).numpy()
    bn = bias = None
    if use_bn:
        bn = [(1 + 0.1 * torch.randn(cout, generator=g)).numpy(), (0.1 * torch.randn(cout, generator=g)).numpy(),
              (0.1 * torch.randn(cout, generator=g)).numpy(), (0.5 + torch.rand(cout, generator=g)).numpy()]
        if cin == 3:
            bias = (0.1 * torch.randn(cout, generator=g)).numpy()
    else:
        bias = (0.1 * torch.randn(cout, generator=g)).numpy()
    res = None
    if res_mode:
        pad = (k - 1) // 2
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        res = torch.randn(N, Ho, Wo, cout, generator=g).half().to(DEV)
    y = conv2d_abi(x, w, stride, bn, bias, act, res, res_mode, out_f32).float()
    ref = conv2d_ref(x, w, stride, bn, bias, act, res, res_mode)
    assert y.shape == ref.shape
    err = (y - ref).abs()
    tol = (1e-3 if out_f32 else 2e-3) * torch.clamp(ref.abs(), min=1.0)
    bad = (err > tol).sum().item()
    assert bad == 0, f"{name}: {bad} / {err.numel()} elements out of tolerance, max err {err.max().item():.4g}"
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("cps", [1, 2, 3])
@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("deep_k_wide_n", "res_after_act", "1x1_s1_768_concat_k", "head_255_f32")],
                         ids=lambda c: c[0])
def test_conv_split_k(case, cps, monkeypatch):
    """The same convolutions with the K split forced (YDST_FORCE_CPS channel blocks per split): partial sums meet in the
    workspace and are added in split order, so the result must stay within the same tolerance AND be run-to-run identical."""
    monkeypatch.setenv("YDST_FORCE_CPS", str(cps))
    test_conv_parity(case)
    name, N, H, W, cin, cout, k, stride, use_bn, act, res_mode, out_f32 = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, cin, generator=g).half().to(DEV)
    w = (torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))).numpy()
    b = (0.1 * torch.randn(cout, generator=g)).numpy()
    ys = [conv2d_abi(x, w, stride, None, b, act, None, 0, out_f32) for _ in range(3)]
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2]), "split-K reduction must be deterministic"
