#!/usr/bin/env python
"""Run single convolutions through ydst_conv2d with YDST_CONV_TRACE=1 and print the per-phase clocks (GPU diagnostic)."""
import os, sys
os.environ.setdefault("YDST_CONV_TRACE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from util import DEV, conv2d_abi
SH = [(4, 76, 76, 128, 256, 3, 1), (4, 76, 76, 128, 256, 3, 0), (3, 76, 76, 128, 256, 3, 0), (1, 76, 76, 128, 256, 3, 0), (4, 76, 76, 256, 128, 1, 0), (4, 38, 38, 256, 512, 3, 1), (4, 38, 38, 256, 512, 3, 0),
      (4, 152, 152, 64, 128, 3, 1), (208, 64, 32, 64, 64, 3, 2), (208, 32, 16, 128, 128, 3, 2)]
for N, H, W, cin, cout, k, rm in SH:
    x = torch.randn(N, H, W, cin).half().to(DEV)
    w = (torch.randn(cout, cin, k, k) * float(np.sqrt(2.0 / (cin * k * k)))).numpy()
    bn = [np.ones(cout, np.float32), np.zeros(cout, np.float32), np.zeros(cout, np.float32), np.ones(cout, np.float32)]
    res = torch.randn(N, H, W, cout).half().to(DEV) if rm else None
    print(f"== N{N} {H}x{W} {cin}->{cout} k{k} res_mode {rm}", file=sys.stderr, flush=True)
    for _ in range(2):
        conv2d_abi(x, w, 1, bn, None, 1, res, rm, False)
