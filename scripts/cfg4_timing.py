"""BASELINE config 4 (scaled): robots of random shapes {biped, worm, box} w x h, hilly terrain,
uniform-t+vxy+a-0 sensors, PhaseSin random phases; bucketed by shape inside the library."""
import importlib, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import helpers

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ft = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
rng = np.random.default_rng(2)
bodies = {}
robots = []
for i in range(n):
    kind = ("biped", "worm", "box")[rng.integers(3)]
    w = int(rng.integers(4 if kind == "biped" else 2, 11)); h = int(rng.integers(2 if kind == "biped" else 1, 11))
    key = f"{kind}-{w}x{h}"
    if key not in bodies:
        bodies[key] = helpers.body_of(key, "uniform-t+vxy+a-0")
    b = bodies[key]
    ph = rng.uniform(0, 2 * math.pi, size=(b.w, b.h))
    robots.append(H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(b.w, b.h, lambda x, y: float(ph[x, y]))), b))
loco = helpers.locomotion(ft, terrain="hilly-1-10-0")
t0 = time.time(); bd = loco.compile(robots); tc = time.time() - t0
with H.DeviceBatch(bd) as db:
    db.upload(0)
    for _ in range(2):
        t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
    launches = db.last_launches()
st = np.bincount(r["status"], minlength=4)
print("cfg4 mixed shapes: n=%d shapes=%d launches=%d compile %.1fs run %.3fs %.3e voxel-steps/s status counts %s" % (
    n, len(bodies), launches, tc, t1 - t0, bd.voxel_steps() / (t1 - t0), st.tolist()), flush=True)
bad = np.nonzero(r["status"])[0][:10]
for i in bad:
    v = robots[i].get_voxels()
    print("   status", r["status"][i], "robot", i, f"{v.w}x{v.h}", sum(x is not None for x in v.values()), "voxels")
