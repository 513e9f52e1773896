import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
robots = helpers.phase_robots(16, seed=21)
bd = helpers.locomotion(2.0).compile(robots, precision=0, trajectory_robots=16)
res, traj = helpers.gpu_run(bd, trajectories=16)
ores, pr = oracle.run(bd, trajectory=True, stats=True)
ot = pr["trajectory"][:, :, :40]
nc = pr["n_contacts"]
for k in range(0, 30):
    d = np.abs(traj[:, k, :, :2] - ot[:, k, :, :2]).max(axis=(1, 2))
    print(k, "max %.2e argmax robot %d" % (d.max(), d.argmax()), "nc", nc[:, k].tolist())
