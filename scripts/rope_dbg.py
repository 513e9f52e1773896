"""error growth with rope limits: kernel fp64 vs oracle per tick"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
shape, lo, hi = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
terrain = sys.argv[4] if len(sys.argv) > 4 else "hilly-1-10-0"
m = H.abi.default_material(); m.area_ratio_passive_min, m.area_ratio_passive_max = lo, hi
g = H.RobotUtils.build_shape(shape)
body = H.Grid.create(g.w, g.h, lambda x, y: H.Voxel([H.AreaRatio()], m) if g.get(x, y) else None)
rng = np.random.default_rng(33)
robots = [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(g.w, g.h, lambda x, y: float(rng.uniform(0, 6.28)))), body) for _ in range(3)]
bd = helpers.locomotion(3.0, terrain=terrain).compile(robots, precision=1, trajectory_robots=3)
with H.DeviceBatch(bd) as db:
    db.upload(0).run()
    tr = np.stack([db.trajectory(i) for i in range(3)])
_, pr = oracle.run(bd, trajectory=True, stats=True, n_threads=3)
nb = 4 * bd.n_voxels[0]
err = np.abs(tr[:, :, :nb, :3] - pr["trajectory"][:, :, :nb, :3]).max(axis=(2, 3))
for k in range(0, err.shape[1], 6):
    print("tick %3d err %.2e %.2e %.2e contacts %d %d %d pos_iters %s" % (k, err[0, k], err[1, k], err[2, k], pr["n_contacts"][0, k], pr["n_contacts"][1, k], pr["n_contacts"][2, k], pr["pos_iters"][:, k] if "pos_iters" in pr else "?"))
