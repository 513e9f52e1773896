import importlib, os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts"); sys.path.insert(0, "/root/repo/tests")
H = importlib.import_module("2dhmsr_b200")
from gpu_check import phase_batch
for sensors in ("uniform-ax+t-0", "empty"):
    loco, robots = phase_batch(4096, sensors=sensors, final_t=20.0, seed=3)
    bd = loco.compile(robots)
    with H.DeviceBatch(bd) as db:
        db.upload(0)
        for _ in range(3):
            t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
    print(sensors, "%.4fs %.3e" % (t1 - t0, bd.voxel_steps() / (t1 - t0)))
