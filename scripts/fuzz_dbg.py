"""error curve of one randomised world (tests/test_parity_gpu.py::_random_world): kernel fp64 vs oracle per tick"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_parity_gpu as T
from oracle import oracle
seed = int(sys.argv[1])
loco, robots, nv, what = T._random_world(seed)
print(what, "spring_mass_mode", loco.settings.spring_mass_mode, "self_collision", loco.settings.self_collision)
bd, res, tr, ores, otr = T._both(loco, robots, T.F64, self_collision=bool(loco.settings.self_collision))
_, pr = oracle.run(bd, stats=True, trajectory=False, self_collision=bool(loco.settings.self_collision))
nb = 4 * nv
err = np.abs(tr[:, :, :nb, :3] - otr[:, :, :nb, :3]).max(axis=(2, 3))
for r in range(err.shape[0]):
    print("robot", r, "status", res["status"][r], ores["status"][r])
    for k in range(0, err.shape[1], 5):
        print("  tick %3d err %.3e  n_contacts %s pos_iters %s" % (k, err[r, k], pr["n_contacts"][r, k], pr.get("pos_iters", np.zeros((3, 999)))[r, k] if "pos_iters" in pr else ""))
