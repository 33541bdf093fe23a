"""GPU diagnostic (not a test): how far is the CUDA path from the fp32 / half-storage oracles on the bench workload, stage by
stage, and on which frame do the track ids of a clip first differ.   python tools/e2e_diag.py [yolov3|yolov4] [frames] [micro_batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workload as W                                            # noqa: E402
from oracle import darknet_ref as D, reid_ref as R              # noqa: E402
from oracle.clip import ClipOracle                              # noqa: E402
from yolo_deepsort_b200 import Darknet, DeepSort, FramePipeline  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    torch.set_num_threads(os.cpu_count())
    dev = torch.device("cuda:0")
    blocks = D.parse_cfg(os.path.join(ROOT, "config", cfg + ".cfg"))
    _, ws = W.darknet_workload(cfg, 608)
    sd = W.reid_workload()
    scenes = W.scenes(608, 608)
    model = Darknet(os.path.join(ROOT, "config", cfg + ".cfg"), img_size=(608, 608))
    model.set_weights(W.flatten_darknet(ws)); model.to(dev)
    ds = DeepSort(sd, use_cuda=True, device=str(dev), **W.TRACKER_KW)
    pipe = FramePipeline(model, ds, W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"], micro_batch=mb)
    mk = lambda half: ClipOracle(blocks, ws, sd, scenes, W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"], W.TRACKER_KW, half)
    o32, o16 = mk(False), mk(True)
    # ---- per scene: detections and features ----
    probe = FramePipeline(model, DeepSort(sd, use_cuda=True, device=str(dev), **W.TRACKER_KW), W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"])
    for si, f in enumerate(scenes):
        _, dets = probe.step(f)
        for name, o in (("fp32", o32), ("half", o16)):
            ref = o.detect(si)[0]
            if ref.shape != dets.shape:
                print(f"scene {si} vs {name}: {len(dets)} detections vs {len(ref)}"); continue
            db = np.abs(dets[:, :4] - ref[:, :4])
            size = np.maximum(ref[:, 2] - ref[:, 0], ref[:, 3] - ref[:, 1])
            print(f"scene {si} vs {name}: n {len(ref)} classes equal {np.array_equal(dets[:, 5], ref[:, 5])} box max abs {db.max():.4f} px "
                  f"(rel to box size {(db.max(1) / size).max():.2e}) score max err {np.abs(dets[:, 4] - ref[:, 4]).max():.2e} "
                  f"int corners equal {np.array_equal(dets[:, :4].astype(np.int64), ref[:, :4].astype(np.int64))}")
        tl = o32.detect(si)[1]
        fr = o32.features(si).numpy()
        fg = ds.extractor.extract(torch.from_numpy(f).to(dev), torch.from_numpy(tl).to(dev)).cpu().numpy()
        rel = np.linalg.norm(fg - fr, axis=1) / np.linalg.norm(fr, axis=1)
        dr, dg = 1 - fr @ fr.T, 1 - fg @ fg.T
        print(f"scene {si} ReID: rel L2 err max {rel.max():.2e} median {np.median(rel):.2e}; pairwise cosine-distance err max {np.abs(dr - dg).max():.2e}")
    # ---- the clip ----
    first = {"fp32": None, "half": None}
    outs = list(pipe.run([scenes[W.clip_index(t)] for t in range(n_frames)]))
    nrows = 0
    for t in range(n_frames):
        si = W.clip_index(t)
        got = np.asarray(outs[t][0], np.int32).reshape(-1, 6)
        nrows += len(got)
        for name, o in (("fp32", o32), ("half", o16)):
            ref, _ = o.step(si)
            same = ref.shape == got.shape and np.array_equal(ref[:, 4:], got[:, 4:])
            if not same and first[name] is None:
                first[name] = t
                print(f"frame {t} (scene {si}) vs {name}: rows {len(got)} vs {len(ref)}; first id mismatch")
            if same and first[name] is None and len(ref):
                d = np.abs(ref[:, :4] - got[:, :4]).max()
                if d > 1:
                    print(f"frame {t} vs {name}: ids equal, track boxes differ by {d} px")
    print(f"clip of {n_frames} frames at micro-batch {mb}: {nrows} track rows; first id mismatch vs fp32 oracle: {first['fp32']}, vs half oracle: {first['half']}")


if __name__ == "__main__":
    main()
