"""Throughput of the other BASELINE configs (scaled down): MLP worms, hilly mixed shapes."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import helpers

def timed(name, loco, robots):
    t0 = time.time(); bd = loco.compile(robots); tc = time.time() - t0
    with H.DeviceBatch(bd) as db:
        db.upload(0)
        for _ in range(2):
            t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
    d = r["last_center_x"] - r["first_center_x"]
    print("%-40s n=%6d compile %.1fs run %.3fs %.3e voxel-steps/s  status!=0: %d  dist mean %.2f" % (
        name, len(robots), tc, t1 - t0, bd.voxel_steps() / (t1 - t0), int((r["status"] != 0).sum()), d.mean()), flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ft = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
timed("cfg3 worm-8x2 MLP [] ax+t", helpers.locomotion(ft), helpers.mlp_robots(n, hidden=(), seed=1))
timed("cfg3 worm-8x2 MLP [21,21] ax+t", helpers.locomotion(ft), helpers.mlp_robots(n, hidden=(21, 21), seed=1))
timed("worm-8x2 PhaseSin ax+t", helpers.locomotion(ft), helpers.phase_robots(n, shape="worm-8x2", seed=1))
timed("biped-4x3 PhaseSin hilly t+vxy+a", helpers.locomotion(ft, terrain="hilly-1-10-0"), helpers.phase_robots(n, sensors="uniform-t+vxy+a-0", seed=2))
timed("biped-4x3 PhaseSin flat no sensors", helpers.locomotion(ft), helpers.phase_robots(n, sensors="empty", seed=2))
