import os
"""Ad-hoc GPU check: kernel (fp64 / fp32) vs the oracle on small batches + a timing."""
import importlib, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle

def phase_batch(n, shape="biped-4x3", sensors="uniform-ax+t-0", seed=0, terrain="flat", final_t=20.0):
    rng = np.random.default_rng(seed)
    body = H.RobotUtils.build_sensorizing_function(sensors)(H.RobotUtils.build_shape(shape))
    robots = []
    for i in range(n):
        ph = rng.uniform(0, 2 * math.pi, size=(body.w, body.h))
        robots.append(H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(body.w, body.h, lambda x, y: float(ph[x, y]))), body))
    st = H.Settings().set(self_collision=int(os.environ.get("SELF", "1")))
    return H.Locomotion(final_t, H.Locomotion.create_terrain(terrain), st), robots

def main():
    loco, robots = phase_batch(6)
    for prec, name in ((1, "f64"), (0, "f32")):
        bd = loco.compile(robots, precision=prec, trajectory_robots=len(robots))
        with H.DeviceBatch(bd) as db:
            db.upload(0).run()
            res = db.results()
            traj = np.stack([db.trajectory(i) for i in range(len(robots))])
        ores, pr = oracle.run(bd, trajectory=True)
        ot = pr["trajectory"]
        print(name, "status", res["status"], "gpu dist", np.round(res["last_center_x"] - res["first_center_x"], 3),
              "oracle dist", np.round(ores["last_center_x"] - ores["first_center_x"], 3))
        for hz in (1, 10, 25, 60, 120, 240, 600, 1200):
            d = np.abs(traj[:, :hz, :, :3] - ot[:, :hz, :, :3]).max()
            print("   ticks<=%4d max|dpos| %.3e" % (hz, d))
        print("   first centre diff", np.abs(res["first_center_x"] - ores["first_center_x"]).max(),
              np.abs(res["first_center_y"] - ores["first_center_y"]).max())
    # timing
    for n in (4096,):
        loco, robots = phase_batch(n, seed=1)
        bd = loco.compile(robots)
        with H.DeviceBatch(bd) as db:
            db.upload(0)
            for _ in range(2):
                t0 = time.time(); db.run(); res = db.results(); t1 = time.time()
                print("n=%d run %.3fs  %.3e voxel-steps/s  bad=%d" % (n, t1 - t0, bd.voxel_steps() / (t1 - t0), int((res["status"] != 0).sum())))
            d = res["last_center_x"] - res["first_center_x"]
            print("   dist mean %.3f std %.3f min %.3f max %.3f" % (d.mean(), d.std(), d.min(), d.max()))
        t0 = time.time(); ores, _ = oracle.run(bd, n_threads=os.cpu_count(), robot_end=64); t1 = time.time()
        od = ores["last_center_x"] - ores["first_center_x"]
        print("   oracle 64 robots %.2fs (%d thr): dist mean %.3f std %.3f" % (t1 - t0, os.cpu_count(), od.mean(), od.std()))

if __name__ == "__main__":
    main()
