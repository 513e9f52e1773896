"""A small run for compute-sanitizer: a few robots of two shape classes (warp-scope and one-block-per-robot),
ground + intra-robot contacts, a distributed controller with sensors, 1.5 s."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import helpers
loco = helpers.locomotion(1.5, terrain="hilly-1-10-0")
for robots, prec in ((helpers.phase_robots(7, seed=1), 0), (helpers.phase_robots(2, shape="box-6x6", seed=2), 0),
                     (helpers.distributed_robots(3, sensors="spinedTouchSighted-t-t-0.01", signals=2, seed=3), 1),
                     (helpers.mlp_robots(3, shape="biped-4x3", sensors="uniformAll-0.01", hidden=(6,), seed=4), 0)):
    bd = loco.compile(robots, precision=prec, trajectory_robots=1)
    with H.DeviceBatch(bd) as db:
        db.upload(0).run()
        r = db.results()
        t = db.trajectory(0)
    print("status", r["status"].tolist(), "dist", np.round(r["last_center_x"] - r["first_center_x"], 3).tolist(), flush=True)
