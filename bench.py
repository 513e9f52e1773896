#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

metric : voxel-steps/s (sum over robots of voxels x ticks, per second)
value  : configs[1] = 4096 x biped-4x3, random-phase PhaseSin, uniform-ax+t-0 sensors, flat terrain, 20 s
         episode (1200 ticks), default Settings = dyn4j's `new Settings()`, continuous detection ALL included -- per
         GPU (weak scaling), fp32 (the reference computes in double: the `fp64` sub-record is the same workload on the
         fp64 build of the same kernel; `without_continuous_detection` is the same workload with World.solveTOI
         switched off, which is NOT what the reference runs).
step   : one pass of the hot path over the batch = the whole batched episode (one kernel launch per shape class).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--extra all|none]
    torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)

`value`  : device-resident throughput (CUDA events on the launch stream, L2 flushed between steps).
`e2e`    : the same metric through the C ABI with HOST buffers: upload (H2D) + run + results (D2H)
           per step, plus the final NCCL gather of the result records when N > 1.
`extra_configs`: configs[2] (65 536 MLP worms, hidden [] and [21, 21]), configs[3] (65 536 mixed shapes on hilly
           terrain) and the per-GPU share of configs[4] (131 072 bipeds per GPU: `--gpus 8` is the 1 M-robot sweep
           with the NCCL gather), each with its own value / e2e / roofline fraction on its own algorithmic bytes;
           1 warm-up + 2 timed steps each, whatever --steps says.
`--impl reference`: the CPU restatement of the dyn4j path (oracle/, "port": the Java reference
           cannot run in this image), all host threads, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROBOTS = 4096
CONFIG = {
    "workload": "configs[1]: 4096 x biped-4x3, PhaseSin f=1 A=1 phases~U[0,2pi) (numpy PCG64 seed 0), "
                "uniform-ax+t-0 sensors, flat terrain, 20 s = 1200 ticks, default Settings (ground and intra-robot "
                "contacts, continuous detection ALL as dyn4j's `new Settings()`); per GPU; fp32 (the reference's arithmetic is IEEE double: see the fp64 sub-record)",
    "robots_per_gpu": N_ROBOTS, "voxels_per_robot": 10, "ticks": 1200, "precision": "fp32",
    "l2": "flushed between timed steps (256 MiB write)", "parallelism": "robots sharded, no per-step communication",
}
EXTRA_STEPS = 2


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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(n_robots, seed, threads):
    """Time the oracle (C restatement of the dyn4j path) on `n_robots` of configs[1]."""
    import workloads
    from oracle import oracle
    loco, robots = workloads.config2(n_robots, seed)
    bd = loco.compile(robots)
    t0 = time.perf_counter()
    oracle.run(bd, n_threads=threads)
    dt = time.perf_counter() - t0
    return bd.voxel_steps() / dt, dt


def cpu_baseline(budget_s=12.0):
    threads = os.cpu_count() or 1
    _, pilot = cpu_port_run(threads, 123, threads)           # one robot per thread
    n = int(max(threads, min(N_ROBOTS, threads * max(1, round(budget_s / max(pilot, 1e-3))))))
    v, dt = cpu_port_run(n, 0, threads)
    return {"value": v, "unit": "voxel-steps/s", "cores": threads, "value_per_core": v / threads, "kind": "port",
            "sample": f"{n} robots of the same workload (full 1200-tick episodes), one robot per thread, "
                      f"{threads} threads, {dt:.1f} s; fp64 C restatement of the dyn4j path -- not the JVM"}


def reference_arm(args):
    """--impl reference: the CPU port on all host threads; a step = the whole 4096-robot batch when the host gets
    through it in <= 15 s, else a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads
    from oracle import oracle
    threads = os.cpu_count() or 1
    _, pilot = cpu_port_run(threads, 123, threads)
    rate = threads / max(pilot, 1e-3)                       # robots per second
    total_budget = 150.0                                    # the whole --steps K run, seconds
    per_step = int(min(N_ROBOTS, max(threads, rate * min(15.0, total_budget / max(1, args.steps)))))
    per_step = max(threads, per_step // threads * threads)
    for _ in range(min(args.warmup, 2)):
        cpu_port_run(threads, 5, threads)
    t, vs = 0.0, 0
    for k in range(args.steps):
        loco, robots = workloads.config2(per_step, k)
        bd = loco.compile(robots)
        t0 = time.perf_counter()
        oracle.run(bd, n_threads=threads)
        t += time.perf_counter() - t0
        vs += bd.voxel_steps()
    value = vs / t
    what = "the whole 4096-robot batch" if per_step == N_ROBOTS else f"bounded sample of the {N_ROBOTS}-robot batch"
    sample = (f"{per_step} robots x 1200 ticks per step ({what}), one robot per thread, {threads} threads; "
              "fp64 C restatement of the dyn4j path -- not the JVM")
    print(json.dumps({
        "impl": "reference", "metric": "voxel-steps/s", "value": value, "unit": "voxel-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "voxel-steps/s", "cores": threads, "value_per_core": value / threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------ GPU arm
class Runner:
    """One workload on this rank's GPU: device-resident and end-to-end timings of whole-episode steps."""

    def __init__(self, H, sharding, torch, dist, dev, local, world, flush):
        self.H, self.sharding, self.torch, self.dist = H, sharding, torch, dist
        self.dev, self.local, self.world, self.flush = dev, local, world, flush
        self.stream = torch.cuda.Stream(device=dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def _run_timed(self, db):
        torch, stream = self.torch, self.stream
        with torch.cuda.stream(stream):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            db.run(stream.cuda_stream)
            ev1.record(stream)
        return ev0, ev1

    def _finish(self, db, n_total):
        """Read the step's result back: D2H of the Outcome records (N = 1) or the NCCL gather of every rank's
        records straight from the device result buffers (N > 1)."""
        if self.world > 1:
            self.torch.cuda.current_stream(self.dev).wait_stream(self.stream)
            return self.sharding.gather_results(self.sharding.device_records(db, self.dev), n_total)
        return db.results()

    def measure(self, bd, n_total, steps, warmup, separate=True):
        """Returns dict(ms=[...] device-resident per step, e2e_s=[...] per step, launches, h2d, res).
        separate=True: K device-resident steps, then K end-to-end steps (the headline config).
        separate=False: every timed step is an end-to-end step whose kernel is also bracketed by CUDA events."""
        H, torch, sp = self.H, self.torch, self.stream.cuda_stream
        db = H.DeviceBatch(bd)
        try:
            db.upload(self.local, sp)
            for _ in range(warmup):
                db.run(sp)
                db.results()
            ms, e2e_s, launches = [], [], 0
            if separate:
                self.barrier()
                for _ in range(steps):
                    self.flush.fill_(1)                       # evict L2 between timed steps
                    torch.cuda.synchronize(self.dev)
                    ev0, ev1 = self._run_timed(db)
                    ev1.synchronize()
                    ms.append(ev0.elapsed_time(ev1))
                    launches += db.last_launches()
                self.barrier()
                db.upload(self.local, sp)                    # untimed: NCCL communicator / pinned buffers
                db.run(sp)
                self._finish(db, n_total)
            for _ in range(steps):
                self.flush.fill_(1)
                self.barrier()
                t0 = time.perf_counter()
                db.upload(self.local, sp)                    # H2D of every input of the episode (pinned staging)
                ev0, ev1 = self._run_timed(db)
                res = self._finish(db, n_total)
                torch.cuda.synchronize(self.dev)
                e2e_s.append(time.perf_counter() - t0)
                if not separate:
                    ms.append(ev0.elapsed_time(ev1))
                launches += db.last_launches()
            self.barrier()
            h2d = int(db.lib.vsr_batch_upload_bytes(db.h))
        finally:
            db.close()
        return {"ms": ms, "e2e_s": e2e_s, "launches": launches, "h2d": h2d, "res": res}

    def reduce(self, m):
        """Max over ranks of the summed times (the contract), plus the per-rank lists."""
        torch, dist = self.torch, self.dist
        t = torch.tensor([sum(m["ms"]), sum(m["e2e_s"])], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            allr = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(allr, t)
            per = torch.stack(allr).cpu().numpy()
        else:
            per = t.cpu().numpy()[None]
        return float(per[:, 0].max()), float(per[:, 1].max()), per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--robots", type=int, default=N_ROBOTS, help="robots per GPU (default: the BASELINE config)")
    ap.add_argument("--extra", default="all", choices=["all", "none"], help="also run configs[2..4] (extra_configs)")
    ap.add_argument("--extra-scale", type=float, default=1.0, help="scale the extra configs' robot counts (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    import workloads
    H = importlib.import_module("2dhmsr_b200")
    sharding = importlib.import_module("2dhmsr_b200.sharding")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    R = Runner(H, sharding, torch, dist, dev, local, world, flush)
    warmup = max(args.warmup, 3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # ---- the headline: configs[1], every rank its own 4096-robot slice of the population (weak scaling)
    n_total = args.robots * world
    b, e = sharding.shard_range(n_total, rank, world)
    t0 = time.perf_counter()
    loco, robots = workloads.config2(e - b, seed=rank)
    bd = loco.compile(robots)
    t_compile = time.perf_counter() - t0
    vs_local = bd.voxel_steps()
    bytes_vs = workloads.algorithmic_bytes_per_voxel_step(robots[:1])
    sampler = ClockSampler(local)
    sampler.start()
    m = R.measure(bd, n_total, args.steps, warmup, separate=True)
    clocks = sampler.stop()
    res = m["res"]
    assert res.shape[0] == n_total and np.all(res["status"] == 0)
    ms_max, e2e_max, per = R.reduce(m)
    total_vs = vs_local * world
    value = total_vs * args.steps / (ms_max / 1e3)
    e2e_value = total_vs * args.steps / e2e_max
    out = None
    if rank == 0:
        launch_s = (ms_max / 1e3) / args.steps
        achieved = bytes_vs * vs_local / launch_s / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "static: " + tj.get("source", "profiles/traffic.json") + " (an ncu --set full capture; not measured in this run)"
        except Exception:
            pass
        out = {
            "metric": "voxel-steps/s", "value": value, "unit": "voxel-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, robots_per_gpu=args.robots),
            "e2e": {"value": e2e_value, "unit": "voxel-steps/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": int((e - b) * 88)},
            "gpu_launches": m["launches"],
            "clocks": clocks,
            "compile_s": round(t_compile, 3),
            "per_rank": {"ms_per_step": [float(x) / args.steps for x in per[:, 0]],
                         "e2e_ms_per_step": [1e3 * float(x) / args.steps for x in per[:, 1]]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "vsr_episode_kernel<float, PHASE_SIN, MODE 3 (default voxel + continuous detection)>",
                         "algorithmic_bytes_per_voxel_step": bytes_vs, "peak_source": peak_src,
                         "note": "per GPU; the path is FP32-issue/latency bound, not HBM bound (DESIGN.md)"},
        }

    # ---- the same workload on the fp64 build of the kernel (the reference's arithmetic; the build the 1e-8 parity
    #      tests run): 1 warm-up + 2 timed steps
    bd64 = loco.compile(robots, precision=H.abi.VSR_PRECISION_F64)
    m64 = R.measure(bd64, n_total, EXTRA_STEPS, 1, separate=False)
    ms64, e64, _ = R.reduce(m64)
    if rank == 0:
        out["fp64"] = {"value": total_vs * EXTRA_STEPS / (ms64 / 1e3), "e2e": total_vs * EXTRA_STEPS / e64,
                       "ms_per_step": ms64 / EXTRA_STEPS, "unit": "voxel-steps/s", "steps": EXTRA_STEPS,
                       "status_nonzero": int((m64["res"]["status"] != 0).sum()),
                       "note": "same workload and kernel source, real = double (the parity build)"}
    del bd64
    # ---- and under ContinuousDetectionMode.NONE (the MODE 1 kernel; the headline runs dyn4j's default, ALL: the
    #      time-of-impact pass is part of the reference's World.step -- DESIGN section 1)
    loco_off = H.Locomotion(20.0, loco.ground_profile, H.Settings().set(continuous_detection=0))
    bdc = loco_off.compile(robots)
    mc = R.measure(bdc, n_total, EXTRA_STEPS, 1, separate=False)
    msc, ec, _ = R.reduce(mc)
    if rank == 0:
        out["without_continuous_detection"] = {
            "value": total_vs * EXTRA_STEPS / (msc / 1e3), "e2e": total_vs * EXTRA_STEPS / ec,
            "ms_per_step": msc / EXTRA_STEPS, "unit": "voxel-steps/s", "steps": EXTRA_STEPS,
            "status_nonzero": int((mc["res"]["status"] != 0).sum()),
            "note": "same workload, vsr_settings.continuous_detection = 0 (World.solveTOI skipped: NOT the reference's "
                    "default; lowers nothing but the cost -- the population's mean displacement rises by 8 %)"}
    del bdc

    # ---- extra_configs
    if args.extra == "all":
        sc = args.extra_scale
        n3, n4, n5 = int(65536 * sc), int(65536 * sc), int(131072 * sc)
        specs = [
            ("configs[2] hidden=[]", f"{n3} x worm-8x2 per GPU, CentralizedSensing MLP TANH 32->16 (528 weights), "
             "uniform-ax+t-0, flat, 1200 ticks", lambda: workloads.config3(n3, (), seed=1 + 100 * rank), n3),
            ("configs[2] hidden=[21,21]", f"{n3} x worm-8x2 per GPU, CentralizedSensing MLP TANH 32->21->21->16 "
             "(1507 weights), uniform-ax+t-0, flat, 1200 ticks", lambda: workloads.config3(n3, (21, 21), seed=1 + 100 * rank), n3),
            ("configs[3]", f"{n4} mixed biped/worm/box shapes up to 10x10 per GPU (full size is 262 144), PhaseSin, "
             "uniform-t+vxy+a-0, hilly-1-10-0, 1200 ticks; one launch per shape class on 16 side streams",
             lambda: workloads.config4(n4, seed=2 + 100 * rank), n4),
            ("configs[4] per-GPU share", f"{n5} x biped-4x3 per GPU as configs[1] (8 GPUs = the 1 048 576-robot sweep), "
             "NCCL all_gather of the vsr_result records inside e2e", lambda: workloads.config5_share(n5, rank), n5),
        ]
        extras = []
        for name, desc, build, n_local in specs:
            t0 = time.perf_counter()
            lo, rb = build()
            bdx = lo.compile(rb)
            tc = time.perf_counter() - t0
            bvs = workloads.algorithmic_bytes_per_voxel_step(rb)
            vsx = bdx.voxel_steps()
            del rb
            mx = R.measure(bdx, n_local * world, EXTRA_STEPS, 1, separate=False)
            msx, ex, perx = R.reduce(mx)
            tv = torch.tensor([float(vsx)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tv)
            if rank == 0:
                tot = float(tv[0])
                ach = bvs * vsx / ((msx / 1e3) / EXTRA_STEPS) / 1e9
                extras.append({
                    "name": name, "workload": desc, "robots_per_gpu": n_local, "voxel_steps_per_step": tot,
                    "value": tot * EXTRA_STEPS / (msx / 1e3), "ms_per_step": msx / EXTRA_STEPS, "steps": EXTRA_STEPS,
                    "e2e": {"value": tot * EXTRA_STEPS / ex, "unit": "voxel-steps/s", "h2d_bytes_per_step": mx["h2d"],
                            "d2h_bytes_per_step": int(n_local * 88)},
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "algorithmic_bytes_per_voxel_step": bvs},
                    "status_nonzero": int((mx["res"]["status"] != 0).sum()), "gpu_launches": mx["launches"],
                    "host_compile_s": round(tc, 2),
                    "per_rank_ms_per_step": [float(x) / EXTRA_STEPS for x in perx[:, 0]],
                })
            del bdx
        if rank == 0:
            out["extra_configs"] = extras

    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
