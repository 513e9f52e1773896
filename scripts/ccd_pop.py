import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
import workloads
loco, robots = workloads.config2(1024, seed=0)
for prec, name in ((H.abi.VSR_PRECISION_F64, "f64"), (H.abi.VSR_PRECISION_F32, "f32")):
    for ccd in (0, 1):
        l2 = H.Locomotion(20.0, loco.ground_profile, H.Settings().set(continuous_detection=ccd))
        with H.DeviceBatch(l2.compile(robots, precision=prec)) as db:
            db.upload(0).run(); r = db.results()
        d = r["last_center_x"] - r["first_center_x"]
        print(name, "ccd", ccd, "mean %.3f median %.3f status!=0 %d" % (d.mean(), np.median(d), int((r["status"] != 0).sum())), flush=True)
