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
ap.add_argument("--mlp", default=None, help="CentralizedSensing MLP with these hidden layers, e.g. 0 (none) or 21,21")
ap.add_argument("--sensors", default="uniform-ax+t-0")
ap.add_argument("--terrain", default="flat")
a = ap.parse_args()
if a.mlp is not None:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    hidden = tuple(int(q) for q in a.mlp.split(",") if q and q != "0")
    robots = helpers.mlp_robots(a.n, shape=a.shape, sensors=a.sensors, hidden=hidden, seed=1)
    loco = helpers.locomotion(a.final_t, terrain=a.terrain)
else:
    loco, robots = phase_batch(a.n, shape=a.shape, sensors=a.sensors, terrain=a.terrain, final_t=a.final_t, seed=3)
bd = loco.compile(robots)
with H.DeviceBatch(bd) as db:
    db.upload(0)
    for _ in range(a.reps):
        t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
        print("n=%d ticks=%d  %.4fs  %.3e voxel-steps/s" % (a.n, bd.n_ticks, t1 - t0, bd.voxel_steps() / (t1 - t0)))
