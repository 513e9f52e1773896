"""development: per-tick divergence kernel(fp64) vs oracle with continuous detection on, big robots sharing a block"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
shape = sys.argv[1] if len(sys.argv) > 1 else "box-8x8"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 24
robots = helpers.phase_robots(n, shape=shape, seed=seed)
if len(sys.argv) > 4:
    robots = robots[int(sys.argv[4]):int(sys.argv[4]) + 1]
    n = 1
for ccd in (1,):
    loco = helpers.locomotion(3.0, terrain="hilly-1-10-0", continuous_detection=ccd)
    bd = loco.compile(robots, precision=H.abi.VSR_PRECISION_F64, trajectory_robots=n)
    with H.DeviceBatch(bd) as db:
        db.upload(0).run()
        res = db.results()
        tr = [db.trajectory(i) for i in range(n)]
    ores, pr = oracle.run(bd, trajectory=True, n_threads=8)
    for i in range(n):
        nb = 4 * bd.n_voxels[bd.robot_shape[i]]
        d = np.abs(tr[i][:, :nb, :3] - pr["trajectory"][i][:, :nb, :3]).max(axis=(1, 2))
        first = int(np.argmax(d > 1e-9)) if (d > 1e-9).any() else -1
        if os.environ.get("CCD_DBG_SAVE"):
            np.save(os.path.join(ROOT, "gpurun_out", "ccd_dbg_tr%d.npy" % i), tr[i])
        print("ccd", ccd, "robot", i, "max %.2e" % d.max(), "first tick > 1e-9:", first, " ".join("%.1e" % x for x in d[::30]))
