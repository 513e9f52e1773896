import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
from gpu_check import phase_batch
loco, robots = phase_batch(6)
bd = loco.compile(robots, precision=0, trajectory_robots=6)
with H.DeviceBatch(bd) as db:
    db.upload(0).run(); res = db.results()
    traj = np.stack([db.trajectory(i) for i in range(6)])
ores, pr = oracle.run(bd, trajectory=True, stats=True)
ot = pr["trajectory"]
names = "x y th vx vy w".split()
for k in list(range(0, 40)) + [50, 60, 80, 100, 120]:
    d = np.abs(traj[:, k] - ot[:, k]).max(axis=(0, 1))
    print(k, " ".join("%s %.2e" % (n, v) for n, v in zip(names, d)), "nc", pr["n_contacts"][:, k])
