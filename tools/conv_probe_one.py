#!/usr/bin/env python
"""One convolution shape, a few launches: the target of an `ncu --set full` capture (GPU diagnostic)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from util import DEV, conv2d_abi
N, H, W, cin, cout, k, rm = [int(v) for v in sys.argv[1:8]]
x = torch.randn(N, H, W, cin).half().to(DEV)
w = (torch.randn(cout, cin, k, k) * float(np.sqrt(2.0 / (cin * k * k)))).numpy()
bn = [np.ones(cout, np.float32), np.zeros(cout, np.float32), np.zeros(cout, np.float32), np.ones(cout, np.float32)]
res = torch.randn(N, H, W, cout).half().to(DEV) if rm else None
for _ in range(3):
    conv2d_abi(x, w, 1, bn, None, 1, res, rm, False)
torch.cuda.synchronize()
