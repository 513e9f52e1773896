import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
shape = sys.argv[1] if len(sys.argv) > 1 else "box-10x10"
robots = helpers.phase_robots(2, shape=shape, sensors=(sys.argv[3] if len(sys.argv) > 3 else "uniform-t-0"), seed=17)
bd = helpers.locomotion(float(sys.argv[2]) if len(sys.argv) > 2 else 1.0).compile(robots, precision=1, trajectory_robots=2)
res, tr = helpers.gpu_run(bd, trajectories=2)
ores, pr = oracle.run(bd, trajectory=True, stats=True)
otr = pr["trajectory"]
nb = 4 * bd.n_voxels[0]
for r in range(2):
    d = np.abs(tr[r, :, :nb] - otr[r, :, :nb]).max(axis=2)   # [tick][body]
    bad = np.argwhere(d > 1e-9)
    if len(bad) == 0:
        print("robot", r, "ok", d.max()); continue
    t0 = bad[:, 0].min()
    print("robot", r, "first bad tick", t0, "bodies", sorted(set(bad[bad[:, 0] == t0][:, 1]))[:40], "nc", pr["n_contacts"][r, max(0,t0-1):t0+2])
    for b in sorted(set(bad[bad[:, 0] == t0][:, 1]))[:4]:
        print("   body", b, "voxel", b // 4, "gpu", tr[r, t0, b], "orc", otr[r, t0, b])
