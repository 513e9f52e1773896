"""config 2 with and without vsr_settings.continuous_detection: wall time of the run + the shift of the population"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
import workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
loco, robots = workloads.config2(n, seed=0)
out = {}
for ccd in (0, 1):
    l2 = H.Locomotion(20.0, loco.ground_profile, H.Settings().set(continuous_detection=ccd))
    bd = l2.compile(robots)
    with H.DeviceBatch(bd) as db:
        db.upload(0)
        for _ in range(3):
            t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
    d = r["last_center_x"] - r["first_center_x"]
    out[ccd] = d
    print("continuous_detection=%d: %.4f s  %.3e voxel-steps/s  status!=0 %d  mean d %.3f median %.3f" % (
        ccd, t1 - t0, bd.voxel_steps() / (t1 - t0), int((r["status"] != 0).sum()), d.mean(), np.median(d)), flush=True)

import ctypes
lib = H.abi.lib if hasattr(H.abi, "lib") else None
try:
    c = (ctypes.c_ulonglong * 7)()
    ctypes.CDLL(os.path.join(ROOT, "2dhmsr_b200", "libvsr_b200.so")).vsr_dev_counters(c)
    print("ccd candidates %d  cycles inside the pass %d (%.0f per candidate)  distance evaluations %d  polygons by advances <3 / 3-7 / 8-19 / >=20: %s  (3 runs)" % (c[0], c[1], c[1] / max(1, c[0]), c[2], list(c[3:7])))
except AttributeError:
    pass
