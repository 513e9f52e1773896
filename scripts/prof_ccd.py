import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
import workloads
loco, robots = workloads.config2(4096, seed=0)
l2 = H.Locomotion(2.0, loco.ground_profile, H.Settings().set(continuous_detection=1))
bd = l2.compile(robots)
with H.DeviceBatch(bd) as db:
    db.upload(0)
    for _ in range(2):
        db.run(); r = db.results()
print("ok", int((r["status"] != 0).sum()))
