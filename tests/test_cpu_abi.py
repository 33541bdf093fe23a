"""CPU tier: the C-ABI library loads, exports every symbol include/ydst.h declares (no compute calls without a GPU), the
Python binding table covers the header, the product package never imports the oracle, and the conv planner's choices for the
BASELINE layer shapes stay sane."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from yolo_deepsort_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ydst.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ydst_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def cdll():
    from yolo_deepsort_b200 import build
    build.build()
    import torch  # noqa: F401  (makes libcudart resolvable)
    return ctypes.CDLL(_lib.LIB_PATH)


def test_library_exports_every_declared_symbol(cdll):
    syms = header_symbols()
    assert len(syms) >= 40
    for name in syms:
        assert getattr(cdll, name) is not None, name
    assert cdll.ydst_version() >= 1


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo_deepsort_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    for f in ("workload.py",):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, f)).read(), flags=re.M)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.YdstError):
        _lib.require_cuda()


def tiling(cdll, N, H, W, cin, cout, k):
    f = cdll.ydst_conv_tiling
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_int)] * 4 + [ctypes.POINTER(ctypes.c_double)]
    bn, ks, occ, ctas = [ctypes.c_int() for _ in range(4)]
    us = ctypes.c_double()
    assert f(N, H, W, cin, cout, k, bn, ks, occ, ctas, us) == 0
    return bn.value, ks.value, occ.value, ctas.value, us.value


@pytest.mark.parametrize("shape", [(1, 76, 76, 128, 256, 3), (1, 38, 38, 256, 512, 3), (1, 19, 19, 512, 1024, 3), (1, 19, 19, 1024, 512, 1),
                                   (1, 76, 76, 256, 128, 1), (50, 64, 32, 64, 64, 3), (50, 8, 4, 512, 512, 3), (4096, 64, 32, 64, 64, 3)])
def test_planner_choices(cdll, shape):
    N, H, W, cin, cout, k = shape
    bn, ks, occ, ctas, us = tiling(cdll, *shape)
    assert bn in (32, 64, 128, 256) and bn <= max(32, 1 << (cout - 1).bit_length())
    assert 1 <= ks <= min(16, cin // 64) and occ in (1, 2)
    m_tiles = (N * (H + 2) * (W + 2) + 127) // 128
    tiles = m_tiles * ((cout + bn - 1) // bn) * ks
    # weight-stationary layers with many tiles run the persistent loop: one CTA per SM strides over the tiles
    assert ctas == tiles or (ctas == 148 and tiles >= 6 * 148 and ks == 1 and occ == 1)
    assert 0 < us < 1e5
    if m_tiles < 148:                      # batch-1 detector layers: do not leave most of the 148 SMs idle
        assert ctas >= 16
    assert cdll.ydst_conv_tiling(1, 19, 19, 48, 64, 3, None, None, None, None, None) != 0     # Cin must be a multiple of 64


def test_planner_micro_batch_choices(cdll):
    """Choices the measurements in DESIGN.md 4.1 (7-8) rest on: a layer with more tiles than SMs but fewer than two per SM is
    planned as ONE resident wave of 128 x 256 tiles (two CTAs per SM), and a narrow layer takes a 64-wide N tile so that it gets
    the TMA-store epilogue."""
    bn, ks, occ, ctas, _ = tiling(cdll, 4, 76, 76, 128, 256, 3)          # 191 tiles on 148 SMs
    assert (bn, ks, occ, ctas) == (256, 1, 2, 191)
    bn, ks, occ, ctas, _ = tiling(cdll, 8, 304, 304, 64, 32, 1)          # cout 32: zero-filled weight rows, clipped store
    assert (bn, ks) == (64, 1) and ctas in (148, (8 * 306 * 306 + 127) // 128)
    # ReID layer1 at a micro-batch's worth of crops: 72 KB of weights stay in shared memory, the CTA walks ~48 tiles
    bn, ks, occ, ctas, _ = tiling(cdll, 408, 64, 32, 64, 64, 3)
    assert (bn, ks, occ, ctas) == (64, 1, 1, 148)
    # few waves with streamed weights stay on plain launches (two co-resident CTAs per SM)
    bn, ks, occ, ctas, _ = tiling(cdll, 8, 76, 76, 256, 128, 1)
    assert ctas == (8 * 78 * 78 + 127) // 128 * (128 // bn) and occ == 2


def test_planner_round2_rules(cdll, monkeypatch):
    """The measured rules of DESIGN.md 4.1 (9-12): persistent CTA pairs for 3x3 layers from two tiles per SM on, plain launches
    below, and the knobs that switch both families off."""
    # 3x3 128->256 @76x76 at micro-batch 8: 381 tiles -> 74 clusters of two persistent CTAs (148 CTAs), one CTA per SM
    bn, ks, occ, ctas, _ = tiling(cdll, 8, 76, 76, 128, 256, 3)
    assert (bn, ks, occ, ctas) == (256, 1, 1, 148)
    # ReID layer4 at 368 crops: 348 tiles >= 2 x 148 -> persistent pairs as well
    bn, ks, occ, ctas, _ = tiling(cdll, 368, 8, 4, 512, 512, 3)
    assert (bn, ks, ctas) == (256, 1, 148)
    # 38x38 256->512 (200 tiles: between one and two waves) and 19x19 512->1024 (112 tiles of 128 x 256): plain launches
    bn, ks, occ, ctas, _ = tiling(cdll, 8, 38, 38, 256, 512, 3)
    assert (bn, ks, ctas) == (256, 1, 200)
    bn, ks, occ, ctas, _ = tiling(cdll, 8, 19, 19, 512, 1024, 3)
    assert (bn, ks, ctas) == (128, 1, 28 * 8)                            # paired 128 x 128 tiles, two per SM
    # knobs: no pairs, no persistent loop -> every layer is one tile per CTA again
    monkeypatch.setenv("YDST_CTA2", "0")
    monkeypatch.setenv("YDST_PERSISTENT", "0")
    for shape in [(8, 76, 76, 128, 256, 3), (408, 64, 32, 64, 64, 3), (8, 304, 304, 64, 32, 1)]:
        bn, ks, occ, ctas, _ = tiling(cdll, *shape)
        N, H, W, cin, cout, k = shape
        assert ctas == (N * (H + 2) * (W + 2) + 127) // 128 * ((max(cout, 64) + bn - 1) // bn) * ks
