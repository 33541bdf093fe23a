#!/usr/bin/env python
"""bench.py -- end-to-end frames/s of the detect-and-track hot path on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path (libydst, sm_100a)
    python bench.py --impl reference [...]                        the reference's CPU path (oracle port) on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      one video stream per GPU

One "step" = one 608x608 frame through the whole per-frame path: u8 frame -> Darknet(yolov3) -> YOLO decode -> NMS ->
box hand-off -> crop + cv2-exact resize -> ReID CNN -> Kalman predict -> cosine+Mahalanobis cost -> LSAP -> IoU cost ->
LSAP -> Kalman update -> track bookkeeping -> (K,6) int32 rows on the host.  ~50 detections per frame (workload.py).

Timed legs (per rank; max over ranks; rank 0 prints ONE JSON line):
  value  K steps with the frames already resident in HBM (ydst_pipeline_step_dev); CUDA events on the launching stream.
  e2e    K steps through the reference-facing call with HOST frames (FramePipeline.step -> ydst_pipeline_step): the
         1.1 MB host->device copy of each frame and the device->host read of the track rows are inside the timed region.
Then (untimed): a per-op CUDA-event pass for the roofline object, and -- rank 0, N=1 only -- the CPU baseline.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "end-to-end FPS (608x608, ~50 dets/frame)"
UNIT = "frames/s"
CFG, SIZE = "yolov3", 608
MICRO_BATCH = 8          # consecutive frames per Darknet / ReID forward (3 slots in flight: 24 frames of look-ahead)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------------------------------
# multi-GPU plumbing: independent streams, one per rank; the only collectives are the barrier and the max-reduce of the
# measured time (no data-path collective: track state is per stream)
# ---------------------------------------------------------------------------------------------------------------------
def dist_init(backend):
    import torch.distributed as dist
    world = env_int("WORLD_SIZE", 1)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=env_int("RANK", 0), world_size=world)
    return env_int("RANK", 0), world


def barrier(device=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        if device is not None and device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()


def max_over_ranks(x, device=None):
    """max of a python float over all ranks (identity at world size 1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_fps(steps_this_rank, elapsed_ms_this_rank, device=None):
    """whole-job frames/s = frames processed by ALL ranks / slowest rank's time."""
    total = sum_over_ranks(steps_this_rank, device)
    worst_ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (worst_ms / 1e3), worst_ms


# ---------------------------------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed regions)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self._stop = threading.Event()
        self._thr = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                   # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:                                    # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": round(max(self.power), 1) if self.power else None}


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def build_pipeline(device, micro_batch=1):
    import workload as W
    from yolo_deepsort_b200 import Darknet, DeepSort, FramePipeline
    defs, ws = W.darknet_workload(CFG, SIZE)
    model = Darknet(os.path.join(ROOT, "config", CFG + ".cfg"), img_size=(SIZE, SIZE))
    model.set_weights(W.flatten_darknet(ws))
    model.to(device)
    ds = DeepSort(W.reid_workload(), use_cuda=True, device=str(device), **W.TRACKER_KW)
    pipe = FramePipeline(model, ds, W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"], micro_batch=micro_batch)
    return model, ds, pipe


def profile_ops(pipe, frames_dev, t0, n=3):
    """Per-op CUDA-event timings of the layer graphs over `n` micro-batches (ydst_profile_begin/_end); eager launches with an
    event pair around every op, so nothing overlaps and the durations are per kernel."""
    import workload as W
    from yolo_deepsort_b200._lib import check, lib
    L = lib()
    check(L.ydst_profile_begin())
    B = pipe.micro_batch
    for i in range(n):                                   # one unit = one full micro-batch: B frames submitted, B collected
        for b in range(B):
            pipe.submit(frames_dev[W.clip_index(t0 + i * B + b)], want_dets=False)
        for b in range(B):
            pipe.collect(want_dets=False)
    cap = 4096
    kind, layer = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    flops, nbytes, ms = np.zeros(cap, np.float64), np.zeros(cap, np.float64), np.zeros(cap, np.float32)
    cnt = ctypes.c_int()
    check(L.ydst_profile_end(cap, kind.ctypes.data, layer.ctypes.data, flops.ctypes.data, nbytes.ctypes.data, ms.ctypes.data, ctypes.byref(cnt)))
    k = cnt.value
    return kind[:k], layer[:k], flops[:k], nbytes[:k], ms[:k].astype(np.float64), n


def stage_times(model, ds, frames_dev, dets, device, micro_batch=1, reps=20):
    """Untimed-leg breakdown: ms per call of the detector forward, NMS, and ReID extraction (same boxes as the last frame),
    each looped back to back on the stream between two CUDA events.  The tracker's share is the step time minus these."""
    import torch
    from yolo_deepsort_b200._lib import check, lib, ptr, stream_ptr
    L = lib()
    h = model.handle(micro_batch)
    frames_b = torch.stack([frames_dev[i % len(frames_dev)] for i in range(micro_batch)]).contiguous()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / reps, 4)

    out["detector_forward_batch%d" % micro_batch] = timed(lambda: check(L.ydst_detector_forward_u8(h, ptr(frames_b), None, stream_ptr())))
    dd = torch.zeros((300, 6), device=device); nn = torch.zeros(1, dtype=torch.int32, device=device)
    out["nms"] = timed(lambda: check(L.ydst_detector_nms(h, 0.5, 0.4, ptr(dd), ptr(nn), stream_ptr())))
    if dets is not None and len(dets):
        d = torch.from_numpy(dets[:, :4].copy()).to(device)
        tlwh = torch.stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]], 1).contiguous()
        tl = tlwh.repeat(micro_batch, 1).contiguous()                     # one ReID forward serves a whole micro-batch
        out["reid_extract_m%d" % len(tl)] = timed(lambda: ds.extractor.extract(frames_dev[0], tl))
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic():
    """DRAM bytes per conv launch from the committed ncu --set full capture, if one has been summarised."""
    p = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def run_ours(args):
    import torch
    import workload as W
    from yolo_deepsort_b200._lib import lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (B200); there is no CPU fallback"
    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = dist_init("nccl")
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    K, Wm = args.steps, args.warmup

    model, ds, pipe = build_pipeline(device, args.micro_batch)
    scenes = W.scenes(SIZE, SIZE)
    host = [torch.from_numpy(s).pin_memory() for s in scenes]            # pinned host frames (e2e leg)
    host_np = [h.numpy() for h in host]
    dev = [h.to(device) for h in host]                                   # resident frames (value leg)
    L = lib()
    t = 0
    n_dets, n_trk = [], []
    for _ in range(max(Wm, 3)):
        tracks, dets = pipe.step(dev[W.clip_index(t)]); t += 1
    # ... and two full micro-batches through the look-ahead path, so that the full-batch plans and CUDA graphs exist before the clock starts
    if args.micro_batch > 1:
        sub = col = 0
        while col < 2 * args.micro_batch:
            while sub < 2 * args.micro_batch and pipe.can_submit():
                pipe.submit(dev[W.clip_index(t)]); t += 1; sub += 1
            pipe.collect(); col += 1
    torch.cuda.synchronize()

    clocks = ClockSampler(local)
    # ---------------- value leg: inputs resident in HBM ----------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(device); torch.cuda.synchronize()
    clocks.start()
    launches0 = L.ydst_launch_count()
    # software-pipelined steady state (ydst_pipeline_submit/_collect): the detector half of frame t+1 is enqueued before the
    # ReID + association half of frame t is collected; exactly K frames are submitted AND collected inside the timed region
    e0.record()
    sub = col = 0
    while col < K:
        while sub < K and pipe.can_submit():
            pipe.submit(dev[W.clip_index(t)]); t += 1; sub += 1
        tracks, dets = pipe.collect(); col += 1
        n_dets.append(len(dets)); n_trk.append(0 if tracks is None else len(tracks))
    e1.record()
    torch.cuda.synchronize()
    launches = L.ydst_launch_count() - launches0
    barrier(device)
    ms_value = e0.elapsed_time(e1)
    # ---------------- e2e leg: host frames through the reference-facing call ----------------
    d2h = 0
    barrier(device); torch.cuda.synchronize()
    e0.record()
    sub = col = 0
    while col < K:
        while sub < K and pipe.can_submit():
            pipe.submit(host_np[W.clip_index(t)]); t += 1; sub += 1       # pinned host frame -> async H2D inside the timed region
        tracks, dets = pipe.collect(); col += 1
        d2h += (0 if tracks is None else np.asarray(tracks).nbytes) + dets.nbytes + 32      # rows + detections + counters
    e1.record()
    torch.cuda.synchronize()
    clocks.stop()
    barrier(device)
    ms_e2e = e0.elapsed_time(e1)

    fps, worst_ms = aggregate_fps(K, ms_value, device)
    fps_e2e, worst_e2e = aggregate_fps(K, ms_e2e, device)

    # ---------------- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), untimed pass ----------------
    kind, layer, flops, nbytes, ms, nunits = profile_ops(pipe, dev, t)
    t += nunits * args.micro_batch
    nprof = nunits * args.micro_batch                                     # frames covered by the profile pass
    if args.dump_ops and rank == 0:
        per = len(kind) // nunits                                       # ops per micro-batch (same op list every time)
        with open(args.dump_ops, "w") as fh:
            fh.write("op,kind,layer,gflop,mbytes,us_avg,tflops\n")
            for i in range(per):
                us = float(np.mean(ms[i::per][:nunits])) * 1e3 if len(kind) == per * nunits else float(ms[i]) * 1e3
                fh.write("%d,%d,%d,%.4f,%.3f,%.2f,%.1f\n" % (i, kind[i], layer[i], flops[i] / 1e9, nbytes[i] / 1e6, us,
                                                            flops[i] / max(us, 1e-3) / 1e6))
    conv = kind == 0
    peaks, peak_src = measured_peaks()
    conv_ms = float(ms[conv].sum()) / nprof
    conv_flops = float(flops[conv].sum()) / nprof
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    roof = {"kernel": "conv_tc2_kernel + conv_tc_kernel (tcgen05 implicit-GEMM convolutions, fp16 in / fp32 accumulate)", "bound": "tensor",
            "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src}); kernel timed inside a long step",
            "traffic": ncu_traffic(),
            "launches_per_micro_batch": int(conv.sum() // nunits), "frames_per_micro_batch": args.micro_batch, "flops_per_step": conv_flops,
            "avg_launch_us": round(conv_ms * nprof * 1e3 / max(1, int(conv.sum())), 2),
            "conv_ms_per_step": round(conv_ms, 4), "all_graph_ops_ms_per_step": round(float(ms.sum()) / nprof, 4),
            "share_of_step": round(conv_ms / (worst_ms / K), 4),
            "share_note": "GPU time of the conv launches per frame / wall time per frame; the three pipeline streams overlap, so the shares of all "
                          "kernels can add up to more than 1",
            "hbm_frac_conv": round(float(nbytes[conv].sum()) / nprof / (conv_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), 4) if conv_ms > 0 else None}

    stages = stage_times(model, ds, dev, dets, device, args.micro_batch)
    out = {"metric": METRIC, "value": round(fps, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
           "ms_per_step": round(worst_ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "fp16", "data": "synthetic",
           "config": {"workload": f"{CFG} {SIZE}x{SIZE} + DeepSort (ReID 128x64 crops), 1 stream per GPU, {args.micro_batch} consecutive frames per forward",
                      "dets_per_frame": round(float(np.mean(n_dets)), 1), "track_rows_per_frame": round(float(np.mean(n_trk)), 1),
                      "clip": f"{W.N_SCENES} synthetic scenes held per workload.SCHEDULE ({W.CLIP_LEN}-frame cycle); seeded random weights, BN statistics and head rows calibrated with decision margins",
                      "tracker": W.TRACKER_KW, "detector": W.DETECT_KW,
                      "l2": "per-step working set (124 MB fp16 weights + ~340 MB activations) exceeds the 126 MB L2; no explicit flush",
                      "pipelining": f"look-ahead: the detector half of the next {args.micro_batch} frame(s) of the stream (one forward) overlaps the "
                                    f"crops+ReID of the previous {args.micro_batch} and the association of the {args.micro_batch} before on three CUDA streams; every one of the K frames is "
                                    "submitted and collected inside the timed region; per-frame results identical to the synchronous step",
                      "micro_batch": args.micro_batch,
                      "parallelism": f"{world} independent streams (no data-path collective)"},
           "e2e": {"value": round(fps_e2e, 2), "unit": UNIT, "ms_per_step": round(worst_e2e / K, 4),
                   "h2d_bytes_per_step": int(SIZE * SIZE * 3), "d2h_bytes_per_step": int(d2h // K)},
           "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof, "stage_ms": stages}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(args.cpu_frames, 2)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline leg: the reference's CPU path, restated in oracle/ (the only place bench.py executes
# oracle code); same frames, same weights, all host threads
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference(n_frames, n_warm):
    import torch
    import workload as W
    from oracle import darknet_ref as D, reid_ref as R, sort_ref as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    blocks = D.parse_cfg(os.path.join(ROOT, "config", CFG + ".cfg"))
    _, ws = W.darknet_workload(CFG, SIZE)
    sd = W.reid_workload()
    scenes = W.scenes(SIZE, SIZE)
    kw = {k: v for k, v in W.TRACKER_KW.items() if k != "min_confidence"}
    trk = S.DeepSortRef(lambda fr, tl: R.extract(sd, fr, tl), **kw)
    stage = {"detect": 0.0, "track": 0.0}
    n_dets = []

    def step(t, timed):
        f = scenes[W.clip_index(t)]
        a = time.perf_counter()
        det = D.detect(blocks, ws, f, (SIZE, SIZE), W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"])
        b = time.perf_counter()
        if det is not None:
            tlwh, conf, cls = D.to_tracker_inputs(det, W.DETECT_KW["class_mask"])
            trk.update(tlwh, conf, f, torch.from_numpy(cls))
        c = time.perf_counter()
        if timed:
            stage["detect"] += b - a; stage["track"] += c - b
            n_dets.append(0 if det is None else len(det))

    for t in range(n_warm):
        step(t, False)
    t0 = time.perf_counter()
    for t in range(n_warm, n_warm + n_frames):
        step(t, True)
    dt = time.perf_counter() - t0
    return {"value": round(n_frames / dt, 4), "unit": UNIT, "cores": int(torch.get_num_threads()), "kind": "port",
            "sample": f"{n_frames} frames of the same clip after {n_warm} warm-up frames (oracle/: torch CPU fp32 convs + restated tracker)",
            "ms_per_frame": round(dt / n_frames * 1e3, 2), "detect_ms": round(stage["detect"] / n_frames * 1e3, 2),
            "track_ms": round(stage["track"] / n_frames * 1e3, 2), "dets_per_frame": round(float(np.mean(n_dets)), 1),
            "host_cores": int(cores)}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    K, Wm = args.steps, max(args.warmup, 1)
    import workload as W
    b = cpu_reference(K, Wm)
    out = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
           "ms_per_step": b["ms_per_frame"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"{CFG} {SIZE}x{SIZE} + DeepSort (ReID 128x64 crops), 1 stream, batch 1, CPU",
                      "dets_per_frame": b["dets_per_frame"], "tracker": W.TRACKER_KW, "detector": W.DETECT_KW,
                      "clip": f"{W.N_SCENES} synthetic scenes held per workload.SCHEDULE ({W.CLIP_LEN}-frame cycle); same frames and weights as the CUDA arm"},
           "cpu_baseline": b,
           "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-frames", type=int, default=16, help="frames timed by the cpu_baseline leg (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--micro-batch", type=int, default=MICRO_BATCH, help="consecutive frames of the stream per Darknet/ReID forward")
    ap.add_argument("--dump-ops", default=None, help="write the per-op CUDA-event timings of the layer graphs to this CSV")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 24 if args.steps is None else args.steps             # one step = one frame, ~0.7 s on 8 host cores
        args.warmup = 2 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 256 if args.steps is None else args.steps
    args.warmup = 16 if args.warmup is None else args.warmup
    return run_ours(args)


if __name__ == "__main__":
    main()
