import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers
shape = sys.argv[1] if len(sys.argv) > 1 else "biped-4x3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
loco = helpers.locomotion(5.0, self_collision=int(os.environ.get("SELF", "1")))
robots = helpers.phase_robots(n, shape=shape, seed=3)
bd = loco.compile(robots, precision=1, trajectory_robots=n)
res, tr = helpers.gpu_run(bd, trajectories=n)
ores, pr = oracle.run(bd, trajectory=True, stats=True)
otr = pr["trajectory"]
nb = 4 * bd.n_voxels[0]
for r in range(n):
    d = np.abs(tr[r, :, :nb, :3] - otr[r, :, :nb, :3]).max(axis=(1, 2))
    bad = np.argmax(d > 1e-9) if (d > 1e-9).any() else -1
    print("robot", r, "first tick >1e-9:", bad, "final", d[-1])
    if bad >= 0:
        for k in range(max(0, bad - 3), min(len(d), bad + 6)):
            dd = np.abs(tr[r, k, :nb] - otr[r, k, :nb])
            b = np.unravel_index(np.argmax(dd), dd.shape)
            print("   tick", k, "max %.3e at body %d comp %d" % (dd.max(), b[0], b[1]), "nc", pr["n_contacts"][r, k], "piters", pr["pos_iters"][r, k])
r = 2
for k in ():
    print("tick", k)
    for b in (2, 3, 10, 11):
        print("   body", b, "gpu", np.array2string(tr[r, k, b], precision=6), "\n           orc", np.array2string(otr[r, k, b], precision=6))
