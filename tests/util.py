"""Shared helpers for the GPU parity tests (call the product through its C ABI via ctypes)."""
import ctypes

import numpy as np
import torch

from yolo_deepsort_b200._lib import check, lib, ptr, stream_ptr

DEV = "cuda:0"


def conv2d_abi(x_nhwc, w, stride, bn=None, bias=None, act=0, res=None, res_mode=0, out_f32=False):
    """x_nhwc: (N,H,W,Cin) fp16 (fp32 if Cin==3) CUDA tensor; w: (Cout,Cin,k,k) fp32 numpy."""
    N, H, W, cin = x_nhwc.shape
    cout, _, k, _ = w.shape
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((N, Ho, Wo, cout), dtype=torch.float32 if out_f32 else torch.float16, device=x_nhwc.device)
    w = np.ascontiguousarray(w, np.float32)
    bn_arr = np.ascontiguousarray(np.concatenate(bn), np.float32) if bn is not None else None
    b_arr = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    check(lib().ydst_conv2d(ptr(x_nhwc.contiguous()), N, H, W, cin, w.ctypes.data, cout, k, stride,
                            bn_arr.ctypes.data if bn_arr is not None else None, b_arr.ctypes.data if b_arr is not None else None,
                            act, ptr(res.contiguous()) if res is not None else None, res_mode, ptr(y), int(out_f32), stream_ptr()))
    return y


def act_torch(x, act):
    import torch.nn.functional as F
    if act == 1:
        return F.leaky_relu(x, 0.1)
    if act == 2:
        return x * torch.tanh(F.softplus(x))
    if act == 3:
        return F.relu(x)
    return x


def conv2d_ref(x_nhwc, w, stride, bn=None, bias=None, act=0, res=None, res_mode=0):
    """fp32 torch reference on the same (fp16-rounded) inputs and weights."""
    import torch.nn.functional as F
    k = w.shape[2]
    x = x_nhwc.float().permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).to(x.device)
    if x_nhwc.shape[3] != 3:
        wt = wt.half().float()
    y = F.conv2d(x, wt, None, stride, (k - 1) // 2)
    if bn is not None:
        g, b, m, v = (torch.from_numpy(np.asarray(t, np.float32)).to(x.device).view(1, -1, 1, 1) for t in bn)
        cb = torch.from_numpy(np.asarray(bias, np.float32)).to(x.device).view(1, -1, 1, 1) if bias is not None else 0
        y = (y + cb - m) / torch.sqrt(v + 1e-5) * g + b
    elif bias is not None:
        y = y + torch.from_numpy(np.asarray(bias, np.float32)).to(x.device).view(1, -1, 1, 1)
    r = res.float().permute(0, 3, 1, 2) if res is not None else None
    if res_mode == 2:
        y = y + r
    y = act_torch(y, act)
    if res_mode == 1:
        y = y + r
    return y.permute(0, 2, 3, 1).contiguous()
