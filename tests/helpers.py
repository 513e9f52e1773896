"""Shared builders for the tests (seeded synthetic inputs of the BASELINE configs)."""
import importlib
import math

import numpy as np

H = importlib.import_module("2dhmsr_b200")


def body_of(shape: str, sensors: str = "uniform-ax+t-0"):
    return H.RobotUtils.build_sensorizing_function(sensors)(H.RobotUtils.build_shape(shape))


def config1_robot():
    """Example.java:41-48: biped-4x3, uniform-ax+t-0, TimeFunctions sin(-2 pi t + pi x / w)."""
    body = body_of("biped-4x3")
    ctrl = H.TimeFunctions(H.Grid.create(
        body.w, body.h, lambda x, y: (lambda t: math.sin(-2 * math.pi * t + math.pi * (x / body.w)))))
    return H.Robot(ctrl, body)


def phase_robots(n, shape="biped-4x3", sensors="uniform-ax+t-0", seed=0, freq=1.0, amp=1.0):
    """BASELINE config 2: PhaseSin f=1, A=1, phases ~ U[0, 2 pi) (numpy PCG64, seeded)."""
    rng = np.random.default_rng(seed)
    body = body_of(shape, sensors)
    out = []
    for _ in range(n):
        ph = rng.uniform(0, 2 * math.pi, size=(body.w, body.h))
        out.append(H.Robot(H.PhaseSin(freq, amp, H.Grid.create(body.w, body.h, lambda x, y: float(ph[x, y]))), body))
    return out


def mlp_robots(n, shape="worm-8x2", sensors="uniform-ax+t-0", hidden=(), seed=1, activation="TANH"):
    """BASELINE config 3: CentralizedSensing + MLP, weights ~ U(-1, 1) (Starter.java:104)."""
    rng = np.random.default_rng(seed)
    body = body_of(shape, sensors)
    nin, nout = H.CentralizedSensing.n_of_inputs(body), H.CentralizedSensing.n_of_outputs(body)
    out = []
    for _ in range(n):
        nw = H.MultiLayerPerceptron.count_weights([nin, *hidden, nout])
        mlp = H.MultiLayerPerceptron(activation, nin, list(hidden), nout, rng.uniform(-1, 1, size=nw))
        out.append(H.Robot(H.CentralizedSensing(body, mlp), body))
    return out


def distributed_robots(n, shape="biped-4x3", sensors="uniform-ax+t-0", signals=1, hidden=(2,), seed=1, activation="TANH",
                       directional=True):
    """Starter.java:176-190: DistributedSensing with one MLP per voxel, weights ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    body = body_of(shape, sensors)
    out = []
    for _ in range(n):
        ds = (H.DistributedSensing if directional else H.DistributedSensingNonDirectional)(body, signals)
        for (x, y), v in body:
            if v is None:
                continue
            nin, nout = ds.n_of_inputs(x, y), ds.n_of_outputs(x, y)
            nw = H.MultiLayerPerceptron.count_weights([nin, *hidden, nout])
            ds.get_functions().set(x, y, H.MultiLayerPerceptron(activation, nin, list(hidden), nout, rng.uniform(-1, 1, size=nw)))
        out.append(H.Robot(ds, body))
    return out


def locomotion(final_t=20.0, terrain="flat", **settings):
    # The body-by-body parity suite runs dyn4j's ContinuousDetectionMode.NONE unless a test asks for the pass: without
    # it kernel and oracle agree to 1e-8..1e-7 over seconds, which is what exposes a wrong constant in a solver; the pass
    # (the library's default, vsr_default_settings) has its own tests in every kernel mode, and the full-size and
    # protocol tests run the default settings.
    settings.setdefault("continuous_detection", 0)
    return H.Locomotion(final_t, H.Locomotion.create_terrain(terrain), H.Settings().set(**settings))


def poly_centres(traj, half=0.525):
    """BehaviorUtils.center over VoxelPoly centres from a [ticks][bodies][6] trajectory."""
    lx = np.array([-1, 1, 1, -1]) * half
    ly = np.array([1, 1, -1, -1]) * half
    k = np.arange(traj.shape[1]) % 4
    c, s = np.cos(traj[:, :, 2]), np.sin(traj[:, :, 2])
    px = traj[:, :, 0] + c * lx[k] - s * ly[k]
    py = traj[:, :, 1] + s * lx[k] + c * ly[k]
    return px.mean(1), py.mean(1)


def gpu_run(bd, device=0, trajectories=0):
    with H.DeviceBatch(bd) as db:
        db.upload(device).run()
        res = db.results()
        traj = [db.trajectory(i) for i in range(trajectories)]
    return res, (np.stack(traj) if traj else None)
