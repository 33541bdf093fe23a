"""bench.py --config reid | assoc: BASELINE.json configs[3] (ReID Extractor only, crop batches 32 -> 4096) and configs[4]
(association stress: 2000 tracks x 2000 detections, galleries of 30 rows = 60 000 gallery rows).  Same contract as bench.py's
headline line: K-step windows repeated to >= --min-seconds, median window, CUDA events on the launching stream, max over ranks,
e2e with host buffers, roofline objects, CPU baseline (the oracle) on a bounded sample.  One JSON line on rank 0.
"""
import ctypes
import json
import os
import time

import numpy as np

import bench as B

ROOT = os.path.dirname(os.path.abspath(__file__))
REID_BATCHES = (32, 128, 512, 2048, 4096)


def _boxes(rng, m):
    return np.stack([rng.uniform(0, 500, m), rng.uniform(0, 440, m), rng.uniform(25, 90, m), rng.uniform(50, 160, m)], 1).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# configs[3]: ReID Extractor only
# ---------------------------------------------------------------------------------------------------------------------
def reid_cpu(n_crops, n_steps, n_warm):
    import torch
    import workload as W
    from oracle import reid_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = W.reid_workload()
    frame = W.scenes(608, 608)[0]
    tl = _boxes(np.random.default_rng(0), n_crops)
    for _ in range(n_warm):
        R.extract(sd, frame, tl)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        R.extract(sd, frame, tl)
    dt = time.perf_counter() - t0
    return {"value": round(n_crops * n_steps / dt, 2), "unit": "crops/s", "cores": int(torch.get_num_threads()), "kind": "port",
            "sample": f"{n_steps} forwards of {n_crops} crops (oracle/: cv2-exact crop + resize, torch CPU fp32 ReID net)",
            "ms_per_step": round(dt / n_steps * 1e3, 2), "host_cores": int(cores)}


def run_reid(args):
    import torch
    import workload as W
    from yolo_deepsort_b200 import Extractor
    from yolo_deepsort_b200._lib import lib
    metric, unit = "ReID Extractor throughput (128x64 crops of a 608x608 frame -> 512-d features)", "crops/s"
    if args.impl == "reference":
        if B.env_int("RANK", 0) != 0:
            return
        K, Wm = (args.steps or 4), max(args.warmup or 1, 1)
        b = reid_cpu(64, K, Wm)
        print(json.dumps({"impl": "reference", "metric": metric, "value": b["value"], "unit": unit, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
                          "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": "ReID Extractor only, 64 crops per step (bounded sample), CPU"},
                          "cpu_baseline": b, "e2e": {"value": b["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (B200); there is no CPU fallback"
    local = B.env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = B.dist_init("nccl")
    K, Wm = (args.steps or 20), max(args.warmup or 3, 3)
    total = 4096
    mine = total // world                                   # the batch is split over the ranks: strong scaling
    ex = Extractor(W.reid_workload(), use_cuda=True, max_batch=max(mine, 32), device=str(device))
    frame_host = torch.from_numpy(W.scenes(608, 608)[0]).pin_memory()
    frame = frame_host.to(device)
    rng = np.random.default_rng(rank)
    flops_per_crop = float(lib().ydst_reid_flops_per_crop())
    peaks, peak_src = B.measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    L = lib()

    def bench_batch(m, seconds):
        tl = torch.from_numpy(_boxes(rng, m)).to(device)
        for _ in range(Wm):
            ex.extract(frame, tl)
        torch.cuda.synchronize()
        wins, _ = B.timed_windows(lambda k: [ex.extract(frame, tl) for _ in range(k)], K, seconds, device)
        med = float(np.median(wins)) / K
        return med, wins, tl

    sweep = {}
    for m in REID_BATCHES:
        if world > 1 or m > mine:
            continue
        med, wins, _ = bench_batch(m, min(args.min_seconds, 0.6))
        tf = flops_per_crop * m / (med * 1e-3) / 1e12
        sweep[str(m)] = {"ms_per_forward": round(med, 4), "crops_per_s": round(m / (med * 1e-3), 1), "tflops": round(tf, 1), "frac": round(tf / peak, 4)}
    clocks = B.ClockSampler(local)
    clocks.start()
    launches0 = L.ydst_launch_count()
    med, wins, tl = bench_batch(mine, args.min_seconds)
    launches = (L.ydst_launch_count() - launches0) / (len(wins) * K + Wm)
    # e2e: host frame + host boxes in, features out to the host, every step
    tl_host = tl.cpu().pin_memory()
    feat_host = torch.empty((mine, 512), dtype=torch.float32).pin_memory()

    def e2e_steps(k):
        for _ in range(k):
            f = frame_host.to(device, non_blocking=True)
            t = tl_host.to(device, non_blocking=True)
            feat_host.copy_(ex.extract(f, t), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps(2)
    wins_e, _ = B.timed_windows(e2e_steps, K, args.min_seconds, device)
    clocks.stop()
    st, value, ms_step = B.window_stats(wins, K, world)
    ste, value_e, ms_step_e = B.window_stats(wins_e, K, world)
    tf = flops_per_crop * mine / (ms_step * 1e-3) / 1e12
    out = {"metric": metric, "value": round(value * mine, 1), "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
           "config": {"workload": f"ReID Extractor only: one forward of {total} crops per step, {mine} per GPU ({world} GPU(s)); crops cut and "
                                  "cv2-exactly resized on the device from a resident 608x608 frame",
                      "crops_per_step": total, "l2": "activations of a 4096-crop forward (several GB) exceed the 126 MB L2"},
           "windows": st,
           "e2e": {"value": round(value_e * mine, 1), "unit": unit, "ms_per_step": round(ms_step_e, 4),
                   "h2d_bytes_per_step": int(608 * 608 * 3 + mine * 16), "d2h_bytes_per_step": int(mine * 2048), "windows": ste},
           "gpu_launches": int(round(launches * K)), "clocks": clocks.summary(),
           "roofline": {"kernel": "conv_tc2_kernel + conv_tc_kernel over the 20-conv ReID net (tcgen05, fp16 in / fp32 accumulate)", "bound": "tensor",
                        "achieved": round(tf, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(tf / peak, 4),
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src})", "traffic": None,
                        "flops_per_crop": flops_per_crop, "note": "whole-forward FLOPs / whole-forward time (crop + stem + pool + L2 norm included)"},
           "batches": sweep}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = reid_cpu(64, 3, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# configs[4]: association stress
# ---------------------------------------------------------------------------------------------------------------------
class Crowd:
    """N objects with constant-velocity boxes and a fixed unit appearance vector each on a 4000 x 4000 plane; every frame each is
    observed with box jitter and appearance noise; `churn` of them are replaced by novel objects per frame (spawn / delete /
    id order are exercised).  Deterministic per seed."""

    def __init__(self, n=2000, seed=0, churn=0.01):
        r = self.rng = np.random.default_rng(seed)
        self.n, self.churn = n, churn
        self.pos = r.uniform(100, 3900, (n, 2)).astype(np.float32)
        self.size = np.stack([r.uniform(20, 60, n), r.uniform(40, 120, n)], 1).astype(np.float32)
        self.vel = r.uniform(-2, 2, (n, 2)).astype(np.float32)
        f = r.standard_normal((n, 512)).astype(np.float32)
        self.feat = f / np.linalg.norm(f, axis=1, keepdims=True)

    def step(self):
        r = self.rng
        k = int(self.n * self.churn)
        if k:
            idx = r.choice(self.n, k, replace=False)
            self.pos[idx] = r.uniform(100, 3900, (k, 2))
            f = r.standard_normal((k, 512)).astype(np.float32)
            self.feat[idx] = f / np.linalg.norm(f, axis=1, keepdims=True)
        self.pos = np.clip(self.pos + self.vel, 50, 3950).astype(np.float32)
        tl = np.concatenate([self.pos - self.size / 2 + r.uniform(-0.5, 0.5, (self.n, 2)), self.size + r.uniform(-0.5, 0.5, (self.n, 2))], 1)
        f = self.feat + 0.05 * r.standard_normal((self.n, 512)).astype(np.float32) / np.sqrt(512)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        order = r.permutation(self.n)
        return tl[order].astype(np.float32), f[order].astype(np.float32), np.zeros(self.n, np.int32)


ASSOC_KW = dict(max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
ASSOC_WARM = 34               # frames before timing: every gallery holds its 30 rows (G = 60 000)


def assoc_frames(n, count, seed=0):
    c = Crowd(n, seed)
    return [c.step() for _ in range(count)]


def assoc_cpu(frames, n_warm, n_timed):
    """The oracle tracker over the same frames: warm-up through the same states (so that its galleries hold 60 000 rows too), then
    n_timed timed updates.  Returns (baseline dict, rows of the timed frames)."""
    import torch
    from oracle import sort_ref as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats = {}
    orc = S.DeepSortRef(lambda fr, tl: feats["f"], max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
    img = np.zeros((4000, 4000, 3), np.uint8)
    rows = []
    t_timed = 0.0
    for t, (tl, ft, cl) in enumerate(frames[:n_warm + n_timed]):
        feats["f"] = torch.from_numpy(ft)
        a = time.perf_counter()
        out = orc.update(tl.copy(), None, img, torch.from_numpy(cl.astype(np.float32)))
        dt = time.perf_counter() - a
        if t >= n_warm:
            t_timed += dt
            rows.append(np.asarray(out, np.int32).reshape(-1, 6))
    n = len(frames[0][0])
    return ({"value": round(n_timed / t_timed, 4), "unit": "updates/s", "cores": int(torch.get_num_threads()), "kind": "port",
             "sample": f"{n_timed} DeepSort.update calls at {n} x {n} after {n_warm} warm-up calls through the same states (oracle/: restated "
                       "tracker, torch CPU fp32, scipy-exact LSAP in C; the gate evaluated in row chunks -- the reference's N x M x M temporary "
                       "would need 32 GB, deep_sort/sort/kalman_filter.py:253)",
             "ms_per_step": round(t_timed / n_timed * 1e3, 1), "host_cores": int(cores)}, rows)


KERNEL_NAMES = {100: "kf_predict_kernel", 101: "normalize_rows_kernel", 102: "fill_i32_kernel", 103: "cosine_min_kernel", 104: "cost_finalize_kernel",
                105: "lsap_kernel", 106: "iou_cost_kernel", 107: "kf_update_kernel", 108: "kf_initiate_kernel", 109: "gallery_append_kernel",
                110: "gather_mean_kernel", 111: "transpose_kernel", 112: "cosine_tc (tcgen05 hi/lo GEMM)", 113: "cost_segmin_kernel", 114: "feat_split_kernel"}


def run_assoc(args):
    import torch
    from yolo_deepsort_b200._lib import check, lib
    from yolo_deepsort_b200.deepsort import TrackerHandle
    n = 2000
    metric, unit = "association updates/s (2000 tracks x 2000 detections, 30-row galleries: 60 000 gallery rows)", "updates/s"
    if args.impl == "reference":
        if B.env_int("RANK", 0) != 0:
            return
        K, Wm = (args.steps or 2), ASSOC_WARM
        b, _ = assoc_cpu(assoc_frames(n, Wm + K), Wm, K)
        print(json.dumps({"impl": "reference", "metric": metric, "value": b["value"], "unit": unit, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
                          "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": "association stress 2000 x 2000, G = 60 000, CPU"},
                          "cpu_baseline": b, "e2e": {"value": b["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (B200); there is no CPU fallback"
    local = B.env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = B.dist_init("nccl")
    K, Wm = (args.steps or 8), ASSOC_WARM
    L = lib()
    max_windows = max(2, 160 // K)                           # the crowd never rewinds: every timed update sees a fresh frame
    n_frames = Wm + 2 + 4 + 2 * (max_windows + 1) * K          # two timed regions, each one calibration window + up to max_windows
    frames = assoc_frames(n, n_frames, seed=rank)
    trk = TrackerHandle(device=str(device), cap_tracks=4096, cap_dets=2048, **{"max_dist": 0.3, "max_iou_distance": 0.7, "max_age": 30, "n_init": 3, "nn_budget": 30})
    dev = [(torch.from_numpy(tl).to(device), torch.from_numpy(ft).to(device)) for tl, ft, _ in frames]
    host = [(torch.from_numpy(tl).pin_memory(), torch.from_numpy(ft).pin_memory()) for tl, ft, _ in frames]
    cls = frames[0][2]
    state = {"t": 0, "rows": []}

    def steps_dev(k):
        for _ in range(k):
            tl, ft = dev[state["t"]]
            state["rows"].append(trk.update(tl, ft, cls)); state["t"] += 1

    def steps_host(k):
        for _ in range(k):
            tl, ft = host[state["t"]]
            state["rows"].append(trk.update(tl.to(device, non_blocking=True), ft.to(device, non_blocking=True), cls)); state["t"] += 1

    steps_dev(Wm)                                            # galleries fill up to their 30 rows
    gpu_rows_after_warm = None
    torch.cuda.synchronize()
    state["rows"] = []
    steps_dev(2)                                             # the two frames the oracle is compared on (parity leg below)
    gpu_rows_after_warm = list(state["rows"])
    clocks = B.ClockSampler(local)
    clocks.start()
    launches0 = L.ydst_launch_count()
    t_before = state["t"]
    wins, _ = B.timed_windows(steps_dev, K, args.min_seconds, device, max_windows=max_windows)
    launches = (L.ydst_launch_count() - launches0) / max(1, state["t"] - t_before)
    wins_e, _ = B.timed_windows(steps_host, K, args.min_seconds, device, max_windows=max_windows)
    clocks.stop()
    st, value, ms_step = B.window_stats(wins, K, world)
    ste, value_e, ms_step_e = B.window_stats(wins_e, K, world)
    tab, _ = trk.table()
    # ---- per-kernel roofline objects: CUDA events around every association kernel of 4 updates ----
    check(L.ydst_profile_begin())
    steps_dev(4)
    cap = 4096
    kind, layer = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    flops, nbytes, ms = np.zeros(cap, np.float64), np.zeros(cap, np.float64), np.zeros(cap, np.float32)
    cnt = ctypes.c_int()
    check(L.ydst_profile_end(cap, kind.ctypes.data, layer.ctypes.data, flops.ctypes.data, nbytes.ctypes.data, ms.ctypes.data, ctypes.byref(cnt)))
    k = cnt.value
    peaks, peak_src = B.measured_peaks()
    hbm = float(peaks["hbm_gbs"])
    kernels = {}
    for kd in sorted(set(kind[:k].tolist())):
        sel = kind[:k] == kd
        t_ms = float(ms[:k][sel].mean())
        by, fl = float(nbytes[:k][sel].mean()), float(flops[:k][sel].mean())
        kernels[KERNEL_NAMES.get(kd, str(kd))] = {"launches_per_update": round(float(sel.sum()) / 4, 2), "avg_us": round(t_ms * 1e3, 1),
                                                  "algorithmic_mb": round(by / 1e6, 2), "gbs": round(by / (t_ms * 1e-3) / 1e9, 1),
                                                  "hbm_frac": round(by / (t_ms * 1e-3) / 1e9 / hbm, 4),
                                                  "tflops": round(fl / (t_ms * 1e-3) / 1e12, 2) if fl else None}
    gpu_ms_per_update = float(ms[:k].sum()) / 4
    dom = max(kernels.items(), key=lambda kv: kv[1]["avg_us"] * kv[1]["launches_per_update"])
    out = {"metric": metric, "value": round(value, 2), "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"association stress: DeepSort tracker.update with {n} detections per frame against ~{len(tab)} live tracks "
                                  "(constant-velocity crowd on a 4000 x 4000 plane, 1 % churn per frame), nn_budget 30, one tracker per GPU",
                      "tracks_alive": int(len(tab)), "gallery_rows": int(sum(min(30, int(r[1])) for r in tab if r[4] == 2)),
                      "rows_per_update": round(float(np.mean([len(r) for r in state["rows"][-8:]])), 1),
                      "l2": "cost matrices (3 x 16 MB) + 123 MB of gallery rows per update; no explicit flush"},
           "windows": st,
           "e2e": {"value": round(value_e, 2), "unit": unit, "ms_per_step": round(ms_step_e, 4),
                   "h2d_bytes_per_step": int(n * (16 + 2048)), "d2h_bytes_per_step": int(np.mean([len(r) for r in state["rows"][-8:]]) * 24), "windows": ste},
           "gpu_launches": int(round(launches * K)), "clocks": clocks.summary(),
           "roofline": {"kernel": dom[0], "bound": "hbm", "achieved": dom[1]["gbs"], "peak": hbm, "unit": "GB/s", "frac": dom[1]["hbm_frac"],
                        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})", "traffic": None,
                        "gpu_ms_per_update_all_kernels": round(gpu_ms_per_update, 3),
                        "note": "dominant kernel by time per update; every association kernel is listed under `kernels` with its algorithmic "
                                "bytes (SURVEY 8d) / CUDA-event time"},
           "kernels": kernels}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b, ref_rows = assoc_cpu(frames, Wm, 2)
        out["cpu_baseline"] = b
        same = all(a.shape == r.shape and np.array_equal(a, r) for a, r in zip(gpu_rows_after_warm, ref_rows))
        out["parity"] = {"frames": 2, "rows_bit_exact": bool(same), "rows": [int(len(r)) for r in ref_rows],
                         "note": f"the (K,6) int32 rows of the two updates after the {Wm} warm-up frames (galleries full: G = 60 000) against the oracle "
                                 "tracker run through the same frames"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main(args):
    return run_reid(args) if args.config == "reid" else run_assoc(args)
