"""Oracle: ReID appearance network (CPU, torch fp32), functional restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows deep_sort/deep/model.py:5-95 (BasicBlock, make_layers, Net with reid=True) and
deep_sort/deep/feature_extractor.py:12-58 (Extractor).  Weights use the reference checkpoint
naming (``torch.load(path)['net_dict']``, feature_extractor.py:16): ``conv.0.*`` stem conv (with
bias), ``conv.1.*`` stem BN, ``layer{1..4}.{0,1}.{conv1,bn1,conv2,bn2}``,
``layer{2..4}.0.downsample.{0,1}``.  The classifier head exists in the checkpoint but is unused
with reid=True (model.py:88-92).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .cv_resize_ref import crops_to_batch

STAGES = ((1, 64, 64, False), (2, 64, 128, True), (3, 128, 256, True), (4, 256, 512, True))


def init_state_dict(seed=0):
    """Seeded synthetic checkpoint with the reference's key names and shapes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, bias=False):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))
        if bias:
            sd[name + ".bias"] = 0.1 * torch.randn(cout, generator=g)

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("conv.0", 64, 3, 3, bias=True); bn("conv.1", 64)
    for li, cin, cout, down in STAGES:
        for bi in range(2):
            c_in = cin if bi == 0 else cout
            p = f"layer{li}.{bi}"
            conv(p + ".conv1", cout, c_in, 3); bn(p + ".bn1", cout)
            conv(p + ".conv2", cout, cout, 3); bn(p + ".bn2", cout)
            if bi == 0 and down:
                conv(p + ".downsample.0", cout, c_in, 1); bn(p + ".downsample.1", cout)
    sd["classifier.0.weight"] = torch.randn(256, 512, generator=g) * 0.05
    sd["classifier.0.bias"] = torch.zeros(256)
    bn("classifier.1", 256)
    sd["classifier.4.weight"] = torch.randn(751, 256, generator=g) * 0.05
    sd["classifier.4.bias"] = torch.zeros(751)
    return sd


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def net_forward(sd, x, half_storage=False):
    """Net.forward, reid=True (deep_sort/deep/model.py:81-92).  x: (m,3,128,64) float32.
    half_storage=True is NOT a reference mode (the Extractor never halves): it rounds the conv weights and every materialised
    activation to fp16 while the arithmetic stays fp32 -- the rounding points of fp16 tensor-core convolutions with a fused fp32
    epilogue -- and is used only to measure the storage-format floor of the feature error (tests/test_precision_floor.py).  The
    stem conv (Cin = 3) keeps fp32 operands, as the CUDA path's hi/lo split does."""
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    h = (lambda t: t.half().float()) if half_storage else (lambda t: t)
    x = torch.as_tensor(x).clone()        # ATen-allocated (aligned) storage: oneDNN's fp32 result depends on it
    with torch.no_grad():
        x = F.conv2d(x, sd["conv.0.weight"], sd["conv.0.bias"], 1, 1)
        x = h(F.relu(_bn(x, sd, "conv.1")))
        x = F.max_pool2d(x, 3, 2, 1)
        for li, cin, cout, down in STAGES:
            for bi in range(2):
                p = f"layer{li}.{bi}"
                s = 2 if (bi == 0 and down) else 1
                y = h(F.relu(_bn(F.conv2d(x, h(sd[p + ".conv1.weight"]), None, s, 1), sd, p + ".bn1")))
                y = _bn(F.conv2d(y, h(sd[p + ".conv2.weight"]), None, 1, 1), sd, p + ".bn2")
                if bi == 0 and down:
                    x = h(_bn(F.conv2d(x, h(sd[p + ".downsample.0.weight"]), None, 2, 0), sd, p + ".downsample.1"))
                x = h(F.relu(x.add(y)))
        x = F.avg_pool2d(x, (8, 4), 1)
        x = x.view(x.size(0), -1)
        x = x.div(x.norm(p=2, dim=1, keepdim=True))
    return x


def extract(sd, frame_rgb_u8, boxes_tlwh):
    """DeepSort._get_features (deep_sort/deep_sort.py:133-146): (m,512) float32 torch tensor."""
    if len(boxes_tlwh) == 0:
        return torch.zeros((0, 512), dtype=torch.float32)
    return net_forward(sd, crops_to_batch(frame_rgb_u8, boxes_tlwh))
