"""error growth of a BreakableVoxel batch: kernel fp64 vs oracle per tick (structure malfunction = rigid joints)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
w, h = int(sys.argv[1]), int(sys.argv[2])
terrain = sys.argv[3] if len(sys.argv) > 3 else "flat"
sens = lambda: [H.SoftNormalization(H.AreaRatio()), H.Malfunction()]
rng = np.random.default_rng(21)
body = H.Grid.create(w, h, lambda x, y: H.BreakableVoxel(sens(), 5 + x + 8 * y, {"STRUCTURE": {"FROZEN"}}, {"TIME": 1.2}, 0.3) if (x + y) % 2 else H.Voxel(sens()))
robots = [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(w, h, lambda x, y: float(rng.uniform(0, 6.28)))), body) for _ in range(2)]
bd = helpers.locomotion(4.0, terrain=terrain).compile(robots, precision=1, trajectory_robots=2)
with H.DeviceBatch(bd) as db:
    db.upload(0).run()
    tr = np.stack([db.trajectory(i) for i in range(2)])
    rd = np.stack([db.sensor_readings(i) for i in range(2)])
_, pr = oracle.run(bd, trajectory=True, readings=True, stats=True, n_threads=2)
nb = 4 * bd.n_voxels[0]
err = np.abs(tr[:, :, :nb, :3] - pr["trajectory"][:, :, :nb, :3]).max(axis=(2, 3))
mal = pr["readings"].reshape(2, err.shape[1], bd.n_voxels[0], 2)[..., 1].sum(-1)
for k in range(0, err.shape[1], 8):
    print("tick %3d err %.2e %.2e  broken voxels %d %d  contacts %d pos_iters %s" % (k, err[0, k], err[1, k], mal[0, k], mal[1, k], pr["n_contacts"][0, k], pr["pos_iters"][0, k] if "pos_iters" in pr else "?"))
