"""Import the UNMODIFIED reference from /root/reference with the three compatibility shims it needs
on this image.  Build-container only: /root/reference does not exist on the GPU box,
so nothing under tests -m gpu / smoke() / bench.py may import this module.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Shims (none of them touches the arithmetic of the hot path):
  * ``torch.solve(B, A)`` was removed from torch; the reference calls it at
    deep_sort/sort/kalman_filter.py:192.  Replaced by the same LU solve, ``torch.linalg.solve(A, B)``.
  * ``matplotlib`` (yolo3/detect/img_detect.py:8,12-13) is not installed: stub modules.
  * ``imutils.video.FileVideoStream`` (yolo3/detect/video_detect.py:12,86) is not installed: a
    synchronous stand-in with the same surface (.stream, .start(), .more(), .read(), transform=).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("YDST_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deep_sort"))


def install():
    import torch
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    if not hasattr(torch, "solve") or getattr(torch.solve, "_ydst_shim", False) is False:
        def solve(B, A):
            return torch.linalg.solve(A, B), None
        solve._ydst_shim = True
        torch.solve = solve
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.ticker"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.NullLocator = object
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    if "imutils" not in sys.modules:
        import cv2

        class FileVideoStream:
            def __init__(self, path, transform=None, queue_size=128):
                self.stream = cv2.VideoCapture(path)
                self.transform = transform
                self._next = None

            def start(self):
                return self

            def _pull(self):
                if self._next is None:
                    ok, frame = self.stream.read()
                    if ok:
                        self._next = self.transform(frame) if self.transform else frame

            def more(self):
                self._pull()
                return self._next is not None

            def read(self):
                self._pull()
                f, self._next = self._next, None
                return f

        im = types.ModuleType("imutils")
        vid = types.ModuleType("imutils.video")
        vid.FileVideoStream = FileVideoStream
        im.video = vid
        sys.modules["imutils"] = im
        sys.modules["imutils.video"] = vid
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
