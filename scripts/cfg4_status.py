"""Config 4 at (a share of) full size: which status bits fire, for which shapes."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ft = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
loco, robots = workloads.config4(n, final_t=ft)
bd = loco.compile(robots)
with H.DeviceBatch(bd) as db:
    db.upload(0)
    t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
print("cfg4 n=%d ticks=%d %.2fs %.3e voxel-steps/s" % (n, bd.n_ticks, t1 - t0, bd.voxel_steps() / (t1 - t0)))
for bit, name in ((1, "nonfinite"), (2, "ground>MAXC"), (4, "pairs>MAX_CAND"), (8, "levels>32")):
    idx = np.nonzero(r["status"] & bit)[0]
    print(f"  bit {bit} ({name}): {len(idx)} robots", [(int(i), f"{robots[i].get_voxels().w}x{robots[i].get_voxels().h}", sum(v is not None for v in robots[i].get_voxels().values())) for i in idx[:12]])
