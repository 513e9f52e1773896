#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : voxel-steps/s (sum over robots of voxels x ticks, per second)
config : configs[1] = 4096 x biped-4x3, random-phase PhaseSin, uniform-ax+t-0 sensors, flat
         terrain, 20 s episode (1200 ticks), default Settings -- per GPU (weak scaling).
step   : one pass of the hot path over the batch = the whole batched episode (one kernel launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)

`value`  : device-resident throughput (CUDA events on the launch stream, L2 flushed between steps).
`e2e`    : the same metric through the C ABI with HOST buffers: upload (H2D) + run + results (D2H)
           per step, plus the final NCCL gather of the result records when N > 1.
`--impl reference`: the CPU restatement of the dyn4j path (oracle/, "port": the Java reference
           cannot run in this image), all host threads, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_VOXEL_STEP = 405.6  # 192 bodies + 144 spring impulses + 57.6 welds + 4 control + 8 sensors
N_ROBOTS = 4096
SHAPE, SENSORS, TERRAIN, FINAL_T = "biped-4x3", "uniform-ax+t-0", "flat", 20.0
CONFIG = {
    "workload": "configs[1]: 4096 x biped-4x3, PhaseSin f=1 A=1 phases~U[0,2pi) (numpy PCG64 seed 0), "
                "uniform-ax+t-0 sensors, flat terrain, 20 s = 1200 ticks, default Settings (ground and intra-robot contacts); per GPU",
    "robots_per_gpu": N_ROBOTS, "voxels_per_robot": 10, "ticks": 1200, "precision": "fp32",
    "l2": "flushed between timed steps (256 MiB write)", "parallelism": "robots sharded, no per-step communication",
}


def build_batch(H, n, seed):
    rng = np.random.default_rng(seed)
    body = H.RobotUtils.build_sensorizing_function(SENSORS)(H.RobotUtils.build_shape(SHAPE))
    robots = []
    for _ in range(n):
        ph = rng.uniform(0, 2 * math.pi, size=(body.w, body.h))
        robots.append(H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(body.w, body.h, lambda x, y: float(ph[x, y]))), body))
    loco = H.Locomotion(FINAL_T, H.Locomotion.create_terrain(TERRAIN), H.Settings())
    return loco, robots


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_run(H, n_robots, seed, threads):
    """Time the oracle (C restatement of the dyn4j path) on `n_robots` of the same workload."""
    from oracle import oracle
    loco, robots = build_batch(H, n_robots, seed)
    bd = loco.compile(robots)
    t0 = time.perf_counter()
    oracle.run(bd, n_threads=threads)
    dt = time.perf_counter() - t0
    return bd.voxel_steps() / dt, dt


def cpu_baseline(H, budget_s=12.0):
    threads = os.cpu_count() or 1
    _, pilot = cpu_port_run(H, threads, 123, threads)           # one robot per thread
    n = int(max(threads, min(4096, threads * max(1, round(budget_s / max(pilot, 1e-3))))))
    v, dt = cpu_port_run(H, n, 0, threads)
    return {"value": v, "unit": "voxel-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n} robots of the same workload (full 1200-tick episodes), one robot per thread, "
                      f"{threads} threads, {dt:.1f} s; fp64 C restatement of the dyn4j path -- not the JVM"}


def reference_arm(args):
    """--impl reference: the CPU port on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H = importlib.import_module("2dhmsr_b200")
    threads = os.cpu_count() or 1
    _, pilot = cpu_port_run(H, threads, 123, threads)
    per_step = int(max(threads, min(N_ROBOTS, threads * max(1, round(8.0 / max(pilot, 1e-3))))))
    for _ in range(args.warmup):
        cpu_port_run(H, threads, 5, threads)
    t, vs = 0.0, 0
    for k in range(args.steps):
        loco, robots = build_batch(H, per_step, k)
        bd = loco.compile(robots)
        from oracle import oracle
        t0 = time.perf_counter()
        oracle.run(bd, n_threads=threads)
        t += time.perf_counter() - t0
        vs += bd.voxel_steps()
    value = vs / t
    sample = (f"{per_step} robots x 1200 ticks per step (bounded sample of the 4096-robot batch), one robot per "
              f"thread, {threads} threads; fp64 C restatement of the dyn4j path -- not the JVM")
    print(json.dumps({
        "impl": "reference", "metric": "voxel-steps/s", "value": value, "unit": "voxel-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "voxel-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--robots", type=int, default=N_ROBOTS, help="robots per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    H = importlib.import_module("2dhmsr_b200")
    sharding = importlib.import_module("2dhmsr_b200.sharding")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n_total = args.robots * world                       # weak scaling: fixed work per GPU
    b, e = sharding.shard_range(n_total, rank, world)
    loco, robots = build_batch(H, e - b, seed=rank)     # every rank its own slice of the population
    t0 = time.perf_counter()
    bd = loco.compile(robots)
    db = H.DeviceBatch(bd)
    t_compile = time.perf_counter() - t0
    voxel_steps_local = bd.voxel_steps()
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    db.upload(local, sp)
    for _ in range(max(args.warmup, 3)):
        db.run(sp)
        db.results()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident: inputs already in HBM, K launches timed with CUDA events on the stream
    barrier()
    launches, ms = 0, 0.0
    for _ in range(args.steps):
        flush.fill_(1)                                   # evict L2 between timed steps
        torch.cuda.synchronize(dev)
        with torch.cuda.stream(stream):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            db.run(sp)
            ev1.record(stream)
        ev1.synchronize()
        ms += ev0.elapsed_time(ev1)
        launches += db.last_launches()
    barrier()
    # ---- end to end through the C ABI: host buffers -> H2D -> run -> D2H (+ final gather)
    def e2e_step():
        db.upload(local, sp)                      # H2D of every input of the episode (pinned staging)
        db.run(sp)
        if world > 1:                             # NCCL all_gather straight from the result buffer
            torch.cuda.current_stream(dev).wait_stream(stream)
            return sharding.gather_results(sharding.device_records(db, dev), n_total)
        return db.results()                       # D2H of the Outcome records

    e2e_step()                                    # untimed: NCCL communicator / pinned buffers
    h2d = None
    t_e2e = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        res = e2e_step()
        torch.cuda.synchronize(dev)
        t_e2e += time.perf_counter() - t0
        launches += db.last_launches()
        h2d = int(db.lib.vsr_batch_upload_bytes(db.h))
    barrier()
    clocks = sampler.stop()
    assert res.shape[0] == n_total and np.all(res["status"] == 0)

    t = torch.tensor([ms, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # max over ranks
    ms_max, e2e_max = float(t[0]), float(t[1])
    total_vs = voxel_steps_local * world
    value = total_vs * args.steps / (ms_max / 1e3)
    e2e_value = total_vs * args.steps / e2e_max
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        launch_s = (ms_max / 1e3) / args.steps
        achieved = BYTES_PER_VOXEL_STEP * voxel_steps_local / launch_s / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        out = {
            "metric": "voxel-steps/s", "value": value, "unit": "voxel-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, robots_per_gpu=args.robots, compile_s=round(t_compile, 3)),
            "e2e": {"value": e2e_value, "unit": "voxel-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int((e - b) * 88)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "vsr_episode_kernel<float, PHASE_SIN, full>",
                         "algorithmic_bytes_per_voxel_step": BYTES_PER_VOXEL_STEP, "peak_source": peak_src,
                         "note": "per GPU; the path is FP32-issue/latency bound, not HBM bound (DESIGN.md)"},
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(H)
        print(json.dumps(out))
    db.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
