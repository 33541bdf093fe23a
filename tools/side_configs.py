#!/usr/bin/env python
"""Timings of the two side configs of BASELINE.json on one GPU (not bench lines; numbers quoted in DESIGN.md):
   configs[3] ReID Extractor only, crop batches 32 -> 4096;   configs[4] association stress, 2000 tracks x 2000 detections."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workload as W
from yolo_deepsort_b200 import Extractor
from yolo_deepsort_b200.deepsort import TrackerHandle

dev = torch.device("cuda", 0)
ex = Extractor(W.reid_workload(), use_cuda=True, max_batch=4096, device="cuda:0")
frame = torch.from_numpy(W.scenes(608, 608)[0]).to(dev)
rng = np.random.default_rng(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for m in (32, 128, 512, 2048, 4096):
    tl = torch.from_numpy(np.stack([rng.uniform(0, 500, m), rng.uniform(0, 440, m), rng.uniform(25, 90, m), rng.uniform(50, 160, m)], 1).astype(np.float32)).to(dev)
    for _ in range(3):
        ex.extract(frame, tl)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        ex.extract(frame, tl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"reid m={m:5d}: {ms:8.3f} ms  {m / ms * 1e3:10.0f} crops/s  {m * 2.2429e9 / (ms * 1e-3) / 1e12:7.1f} TFLOP/s")

n = 2000
trk = TrackerHandle(0.3, 0.7, 30, 3, 30, 4096, 2048, "cuda:0")
def frame_inputs(seed, jitter):
    r = np.random.default_rng(seed)
    base = np.random.default_rng(7)
    tl = np.stack([base.uniform(0, 4000, n), base.uniform(0, 4000, n), base.uniform(20, 60, n), base.uniform(40, 120, n)], 1).astype(np.float32)
    ft = base.standard_normal((n, 512)).astype(np.float32)
    tl[:, :2] += r.normal(0, jitter, (n, 2)).astype(np.float32)
    ft = ft + 0.05 * r.standard_normal((n, 512)).astype(np.float32)
    ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    return torch.from_numpy(tl).to(dev), torch.from_numpy(ft).to(dev)
cls = np.zeros(n, np.int32)
times = []
for t in range(8):
    tl, ft = frame_inputs(t, 1.0)
    torch.cuda.synchronize(); a = time.perf_counter()
    out = trk.update(tl, ft, cls)
    torch.cuda.synchronize(); times.append((time.perf_counter() - a) * 1e3)
    print(f"assoc frame {t}: {times[-1]:8.2f} ms, {len(out)} rows, {len(trk.table()[0])} tracks")
print("assoc 2000x2000 steady state: %.2f ms per update (frames 4-7)" % np.mean(times[4:]))
