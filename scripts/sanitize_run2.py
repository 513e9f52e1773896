"""A small run for compute-sanitizer (round 2): several big robots per block (named barriers), MLP on big robots, rope
limits, BreakableVoxels (all components) with a sensing controller, the readings dump; 1.0 s each."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import helpers
loco = helpers.locomotion(1.0, terrain="hilly-1-10-0")
rope = H.abi.default_material(); rope.area_ratio_passive_min, rope.area_ratio_passive_max = 0.9, 1.1
sens = lambda: [H.SoftNormalization(H.AreaRatio()), H.Malfunction(), H.Average(H.Touch(), 0.25)]
def phase(body, n, seed):
    rng = np.random.default_rng(seed)
    return [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(body.w, body.h, lambda x, y: float(rng.uniform(0, 6.28)))), body) for _ in range(n)]
def mlp(body, n, seed):
    rng = np.random.default_rng(seed)
    nin, nout = H.CentralizedSensing.n_of_inputs(body), H.CentralizedSensing.n_of_outputs(body)
    nw = H.MultiLayerPerceptron.count_weights([nin, 6, nout])
    return [H.Robot(H.CentralizedSensing(body, H.MultiLayerPerceptron("TANH", nin, [6], nout, rng.uniform(-1, 1, size=nw))), body) for _ in range(n)]
brk = lambda w, h: H.Grid.create(w, h, lambda x, y: H.BreakableVoxel(sens(), 3 + x + 7 * y, {"ACTUATOR": {"ZERO", "FROZEN", "RANDOM"}, "SENSORS": {"ZERO", "RANDOM"},
                                                                      "STRUCTURE": {"FROZEN"}}, {"TIME": 0.3, "AREA": 20.0}, 0.2))
cases = [
    ("7 x box-8x8 (4 robots per block, fp32)", phase(helpers.body_of("box-8x8", "uniform-t+vxy+a-0"), 7, 1), 0),
    ("5 x box-10x10 (2 robots per block, fp32)", phase(helpers.body_of("box-10x10", "uniform-t+vxy+a-0"), 5, 2), 0),
    ("3 x box-7x6 MLP (fp64, 2 robots per block)", mlp(helpers.body_of("box-7x6", "uniform-ax+t-0"), 3, 3), 1),
    ("rope limits, biped (fp32)", phase(H.Grid.create(4, 3, lambda x, y: H.Voxel([H.AreaRatio()], rope)), 5, 4), 0),
    ("rope limits, box-6x6 (fp32)", phase(H.Grid.create(6, 6, lambda x, y: H.Voxel([H.AreaRatio()], rope)), 3, 5), 0),
    ("breakable 4x2 MLP (fp32)", mlp(brk(4, 2), 5, 6), 0),
    ("breakable 7x6 MLP (fp32)", mlp(brk(7, 6), 3, 7), 0),
]
for name, robots, prec in cases:
    bd = loco.compile(robots, precision=prec, trajectory_robots=1)
    with H.DeviceBatch(bd) as db:
        db.upload(0).run()
        r = db.results()
        t = db.trajectory(0)
        rd = db.sensor_readings(0)
    print(name, "status", r["status"].tolist(), "dist", np.round(r["last_center_x"] - r["first_center_x"], 3).tolist(), flush=True)
