"""CPU tier: the N>1 plumbing of bench.py (one independent stream per rank, no data-path collective) at world size 2 over
gloo: whole-job frames/s = frames of ALL ranks / slowest rank's time, and only rank 0 of the reference arm does work."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    r, w = bench.dist_init("gloo")
    assert (r, w) == (rank, world)
    bench.barrier()
    # rank 0 processed 10 frames in 100 ms, rank 1 processed 10 frames in 250 ms -> 20 frames / 0.25 s
    fps, worst = bench.aggregate_fps(10, 100.0 if rank == 0 else 250.0)
    per_rank = bench.gather_floats(10.0 + rank)                 # per-rank ms, in rank order (names a straggler)
    q.put((rank, fps, worst, bench.max_over_ranks(float(rank)), bench.sum_over_ranks(1.0), per_rank))
    bench.barrier()
    dist.destroy_process_group()


def test_world_size_2_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fps, worst, mx, sm, per_rank in res:
        assert abs(fps - 80.0) < 1e-9 and worst == 250.0 and mx == 1.0 and sm == 2.0 and per_rank == [10.0, 11.0]


def test_window_statistics():
    """value / ms_per_step come from ALL timed steps over the whole timed region (the windows are contiguous); count, median, min
    and max of the single windows are reported next to it."""
    sys.path.insert(0, ROOT)
    import bench
    st, fps, ms_step = bench.window_stats([10.0, 12.0, 50.0], 20, 2)
    assert st["count"] == 3 and st["ms_per_step_median"] == 0.6 and st["ms_per_step_min"] == 0.5 and st["ms_per_step_max"] == 2.5
    assert st["ms_per_step_mean"] == 1.2 and st["timed_s"] == 0.072
    assert abs(fps - 2 * 3 * 20 / 0.072) < 1e-6 and abs(ms_step - 1.2) < 1e-12


def test_reference_arm_non_zero_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
