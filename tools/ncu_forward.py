#!/usr/bin/env python
"""Target of the `ncu --set full` capture of EVERY kernel of one steady-state micro-batch (GPU diagnostic):
one Darknet forward at the bench's micro-batch, its batched NMS + hand-off, one ReID forward over the crops of those frames and
one DeepSort.update per frame, between cudaProfilerStart / Stop (run ncu with --profile-from-start off; YDST_GRAPH=0 so that
every kernel is its own launch).   python tools/ncu_forward.py [yolov3|yolov4] [micro_batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                            # noqa: E402
import workload as W                                         # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda", 0)
    model, ds, pipe = B.build_pipeline(cfg, dev, mb)
    frames = [torch.from_numpy(s).to(dev) for s in W.scenes(608, 608)]
    t = 0
    for _ in range(3 * mb):                                  # warm-up: plans, tensor maps, galleries
        pipe.submit(frames[W.clip_index(t)]); t += 1
        if not pipe.can_submit():
            for _ in range(mb):
                pipe.collect()
    pipe.drain()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(mb):
        pipe.submit(frames[W.clip_index(t)]); t += 1
    for _ in range(mb):
        pipe.collect()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one micro-batch of", mb, "frames of", cfg)


if __name__ == "__main__":
    main()
