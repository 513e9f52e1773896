import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import helpers
loco = helpers.locomotion(5.0, gravity_y=0.0)
bd = loco.compile(helpers.phase_robots(4096, seed=3))
with H.DeviceBatch(bd) as db:
    db.upload(0)
    for _ in range(2):
        t0 = time.time(); db.run(); r = db.results(); t1 = time.time()
    print("nocontact n=4096 %.4fs %.3e" % (t1 - t0, bd.voxel_steps() / (t1 - t0)))
