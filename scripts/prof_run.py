"""Short run for ncu: N robots, a few seconds of simulated time."""
import argparse, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
H = importlib.import_module("2dhmsr_b200")
from gpu_check import phase_batch
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--final-t", type=float, default=2.0)
ap.add_argument("--shape", default="biped-4x3")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
loco, robots = phase_batch(a.n, shape=a.shape, final_t=a.final_t, seed=3)
bd = loco.compile(robots)
with H.DeviceBatch(bd) as db:
    db.upload(0)
    for _ in range(a.reps):
        t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
        print("n=%d ticks=%d  %.4fs  %.3e voxel-steps/s" % (a.n, bd.n_ticks, t1 - t0, bd.voxel_steps() / (t1 - t0)))
