"""Seeded synthetic populations of the BASELINE.json configs -- shared by bench.py and tests/.

config 2: n x biped-4x3, PhaseSin f=1 A=1, phases ~ U[0, 2 pi), uniform-ax+t-0, flat              (seed 0)
config 3: n x worm-8x2, CentralizedSensing + MLP TANH, hidden [] or [21, 21], weights ~ U(-1, 1)
          (Starter.java:104,143), uniform-ax+t-0, flat                                             (seed 1)
config 4: n robots of shapes drawn uniformly from {biped, worm, box} w x h (2 <= w <= 10, 1 <= h <= 10; biped
          w >= 4, h >= 2), PhaseSin random phases, uniform-t+vxy+a-0, hilly-1-10-0                 (seed 2)
config 5: config 2 at 1 048 576 robots over the GPUs of one box (131 072 per GPU at 8 GPUs)

and the ALGORITHMIC bytes per voxel-step of each (state streamed once in and once out of HBM
per tick, fp32): bodies 192 + spring impulses 144 + welds 24 W/V + control 4 + sensor readings 4 S
(+ MLP weights 4 nW / V when the controller is a per-robot network).
"""
from __future__ import annotations

import importlib
import math

import numpy as np

H = importlib.import_module("2dhmsr_b200")

_BODIES = {}


def body_of(shape: str, sensors: str):
    key = (shape, sensors)
    if key not in _BODIES:
        _BODIES[key] = H.RobotUtils.build_sensorizing_function(sensors)(H.RobotUtils.build_shape(shape))
    return _BODIES[key]


def _phase_robot(rng, body):
    ph = rng.uniform(0, 2 * math.pi, size=(body.w, body.h))
    return H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(body.w, body.h, lambda x, y: float(ph[x, y]))), body)


def config2(n: int = 4096, seed: int = 0):
    """(Locomotion, robots): BASELINE configs[1]."""
    rng = np.random.default_rng(seed)
    body = body_of("biped-4x3", "uniform-ax+t-0")
    robots = [_phase_robot(rng, body) for _ in range(n)]
    return H.Locomotion(20.0, H.Locomotion.create_terrain("flat"), H.Settings()), robots


def config3(n: int = 65536, hidden=(), seed: int = 1, final_t: float = 20.0):
    """(Locomotion, robots): BASELINE configs[2]; hidden = () (528 weights) or (21, 21) (1507 weights)."""
    rng = np.random.default_rng(seed)
    body = body_of("worm-8x2", "uniform-ax+t-0")
    nin, nout = H.CentralizedSensing.n_of_inputs(body), H.CentralizedSensing.n_of_outputs(body)
    nw = H.MultiLayerPerceptron.count_weights([nin, *hidden, nout])
    robots = []
    for _ in range(n):
        mlp = H.MultiLayerPerceptron("TANH", nin, list(hidden), nout, rng.uniform(-1, 1, size=nw))
        robots.append(H.Robot(H.CentralizedSensing(body, mlp), body))
    return H.Locomotion(final_t, H.Locomotion.create_terrain("flat"), H.Settings()), robots


def config4(n: int = 262144, seed: int = 2, final_t: float = 20.0, terrain: str = "hilly-1-10-0",
            sensors: str = "uniform-t+vxy+a-0"):
    """(Locomotion, robots): BASELINE configs[3]."""
    rng = np.random.default_rng(seed)
    robots = []
    for _ in range(n):
        kind = ("biped", "worm", "box")[rng.integers(3)]
        w = int(rng.integers(4 if kind == "biped" else 2, 11))
        h = int(rng.integers(2 if kind == "biped" else 1, 11))
        robots.append(_phase_robot(rng, body_of(f"{kind}-{w}x{h}", sensors)))
    return H.Locomotion(final_t, H.Locomotion.create_terrain(terrain), H.Settings()), robots


def config5_share(n_per_gpu: int = 131072, rank: int = 0):
    """(Locomotion, robots): one GPU's share of BASELINE configs[4] (1 048 576 bipeds over 8 GPUs)."""
    return config2(n_per_gpu, seed=1000 + rank)


def algorithmic_bytes_per_voxel_step(robots) -> float:
    """SURVEY 8(d), voxel-weighted over the batch (all robots run the same number of ticks)."""
    cache = {}
    tot_b, tot_v = 0.0, 0
    for r in robots:
        g = r.get_voxels()
        ctrl = r.get_controller()
        key = (id(g), type(ctrl).__name__, len(getattr(getattr(ctrl, "function", None), "weights", ())))
        if key not in cache:
            cells = {(x, y) for (x, y), v in g if v is not None}
            v_n = len(cells)
            welds = 2 * sum(((x - 1, y) in cells) + ((x, y - 1) in cells) for (x, y) in cells)  # Robot.java:101-110
            s_n = sum(len(s.domains()) for v in g.values() if v is not None for s in v.get_sensors())
            n_w = key[2]
            cache[key] = ((192 + 144 + 4) * v_n + 24.0 * welds + 4.0 * s_n + 4.0 * n_w, v_n)
        b, v_n = cache[key]
        tot_b += b
        tot_v += v_n
    return tot_b / tot_v
