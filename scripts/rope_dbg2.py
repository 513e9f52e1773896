import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
shape = sys.argv[1]
m = H.abi.default_material(); m.area_ratio_passive_min, m.area_ratio_passive_max = 0.9, 1.1
g = H.RobotUtils.build_shape(shape)
body = H.Grid.create(g.w, g.h, lambda x, y: H.Voxel([H.AreaRatio()], m) if g.get(x, y) else None)
rng = np.random.default_rng(33)
robots = [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(g.w, g.h, lambda x, y: float(rng.uniform(0, 6.28)))), body) for _ in range(1)]
bd = helpers.locomotion(0.3, terrain="flat").compile(robots, precision=1, trajectory_robots=1)
with H.DeviceBatch(bd) as db:
    db.upload(0).run()
    tr = np.stack([db.trajectory(i) for i in range(1)])
_, pr = oracle.run(bd, trajectory=True, stats=True)
nb = 4 * bd.n_voxels[0]
e = np.abs(tr[0, :, :nb, :] - pr["trajectory"][0, :, :nb, :])
for k in range(e.shape[0]):
    b = np.unravel_index(np.argmax(e[k]), e[k].shape)
    print("tick %2d max err %.2e at body %d (voxel %d mass %d) comp %d  contacts %d pos_iters %d" % (k, e[k].max(), b[0], b[0] // 4, b[0] % 4, b[1], pr["n_contacts"][0, k], pr["pos_iters"][0, k]))
    if e[k].max() > 1e-9:
        bad = np.argwhere(e[k].max(axis=1) > 1e-9)[:, 0]
        print("   bad bodies:", bad.tolist()[:60])
        break
