"""Host-side mirror vs the reference's own vectors (MLP test, Random comment, DSL semantics)."""
import importlib
import json
import math
import os

import numpy as np
import pytest

H = importlib.import_module("2dhmsr_b200")
A = H.abi
from oracle import oracle  # test infrastructure
import helpers

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_java_random_known_answers():
    g = GOLD["survey_java_random"]
    assert H.JavaRandom(0).next_double() == g["seed0_nextDouble"]
    assert H.JavaRandom(0).next_int() == g["seed0_nextInt"]
    assert H.JavaRandom(42).next_int() == g["seed42_nextInt"]
    assert H.JavaRandom(0).next_gaussian() == pytest.approx(g["seed0_nextGaussian"], abs=1e-15)
    # Locomotion.java:79: "the 1st double of nextDouble() is always around 0.73"
    for seed in range(20):
        assert abs(H.JavaRandom(seed).next_double() - GOLD["reference_random_first_double_about"]) < 0.01
    r = H.JavaRandom(5)
    assert all(0 <= r.next_int(10) < 10 for _ in range(200))
    assert all(0 <= r.next_int(16) < 16 for _ in range(200))


def test_terrain_dsl():
    flat = H.Locomotion.create_terrain("flat")  # Locomotion.java:70-75
    assert flat == [[0.0, 10.0, 1990.0, 2000.0], [100.0, 5.0, 5.0, 100.0]]
    hilly = H.Locomotion.create_terrain("hilly-1-10-0")
    g = GOLD["survey_hilly_1_10_0"]
    assert len(hilly[0]) == g["n_points"]
    for i, (x, y) in enumerate(g["interior"]):
        assert hilly[0][2 + i] == pytest.approx(x, abs=5e-5) and hilly[1][2 + i] == pytest.approx(y, abs=5e-5)
    assert hilly[0][:2] == [0.0, 10.0] and hilly[1][:2] == [100.0, 0.0] and hilly[1][-1] == 100.0
    assert hilly[0][-1] == hilly[0][-2] + 10.0
    assert np.allclose(hilly[0], GOLD["port_hilly_1_10_0"]["xs"]) and np.allclose(hilly[1], GOLD["port_hilly_1_10_0"]["ys"])
    assert all(b - a >= 1.0 - 1e-12 for a, b in zip(hilly[0][1:-2], hilly[0][2:-1]))  # max(1, ...) step
    up = H.Locomotion.create_terrain("uphill-10")
    dy = 1980 * math.sin(math.radians(10))
    assert up[1] == pytest.approx([100.0, 5.0, 5 + dy, 100 + dy])
    down = H.Locomotion.create_terrain("downhill-15.5")
    assert down[1][1] == pytest.approx(5 + 1980 * math.sin(15.5 / 180 * math.pi))
    st = H.Locomotion.create_terrain("steppy-2-8-3")
    assert len(st[0]) == len(st[1]) and len(st[0]) % 2 == 1
    fw = H.Locomotion.create_terrain("flatWithStart-2")
    assert fw[0][2] == 10 + 24.0 and len(fw[0]) == 5
    with pytest.raises(ValueError, match="Unknown terrain name"):
        H.Locomotion.create_terrain("moon")


def test_shapes_and_grid_order():
    b = H.RobotUtils.build_shape("biped-4x3")  # RobotUtils.java:211-215
    assert (b.w, b.h) == (4, 3)
    assert [(x, y) for (x, y), v in b if not v] == [(1, 0), (2, 0)]
    assert sum(bool(v) for v in H.RobotUtils.build_shape("worm-8x2").values()) == 16
    assert sum(bool(v) for v in H.RobotUtils.build_shape("box-10x10").values()) == 100
    tri = H.RobotUtils.build_shape("tripod-5x3")
    assert [(x, y) for (x, y), v in tri if not v] == [(1, 0), (3, 0)]
    comb = H.RobotUtils.build_shape("comb-5x4")
    assert not comb.get(1, 0) and comb.get(1, 2) and comb.get(2, 0)
    assert sum(bool(v) for v in H.RobotUtils.build_shape("ball-5").values()) == 21
    with pytest.raises(ValueError, match="Unknown body name"):
        H.RobotUtils.build_shape("snake-3")
    # Grid.create calls the filler x-outer/y-inner (Grid.java:103-111); values() is row-major (:268-270)
    calls = []
    g = H.Grid.create(2, 3, lambda x, y: calls.append((x, y)) or (x, y))
    assert calls == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2)]
    assert g.values() == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
    assert g.get(-1, 0) is None and g.get(2, 0) is None


def test_sensor_dsl_and_encoding():
    body = helpers.body_of("biped-4x3", "uniform-ax+t-0")
    v = body.get(0, 0)
    assert [type(s).__name__ for s in v.sensors] == ["SoftNormalization", "Average"]
    ax, t = (s.encode() for s in v.sensors)
    assert (ax.base, ax.axes, ax.rotated, ax.aggregator, ax.soft_norm) == (A.VSR_SENSOR_VELOCITY, 1, 1, A.VSR_AGG_TREND, 1)
    assert ax.max_velocity_norm == 4.0 and ax.interval == 0.5
    assert (t.base, t.aggregator, t.soft_norm, t.interval) == (A.VSR_SENSOR_TOUCH, A.VSR_AGG_AVERAGE, 0, 0.25)
    # Trend domain = +- extent / interval (Trend.java:31-33); SoftNormalization -> [0, 1]
    tr = H.Trend(H.Velocity(True, 4.0, "X"), 0.5)
    assert (tr.domains()[0].min, tr.domains()[0].max) == (-16.0, 16.0)
    body4 = helpers.body_of("worm-8x2", "uniform-t+vxy+a-0")
    assert H.CentralizedSensing.n_of_inputs(body4) == 16 * 4 and H.CentralizedSensing.n_of_outputs(body4) == 16
    assert H.CentralizedSensing.n_of_inputs(helpers.body_of("worm-8x2")) == 32
    with pytest.raises(ValueError, match="Unknown sensorizing function name"):
        H.RobotUtils.build_sensorizing_function("uniform-zz-0")
    with pytest.raises(NotImplementedError):
        H.RobotUtils.build_sensorizing_function("spinedTouch-t-f-0.1-x")(H.RobotUtils.build_shape("box-2x2")) \
            if False else H.Lidar(10.0, *([0.0] * 9))       # more than 8 rays per sensor
    l5 = H.RobotUtils.build_sensorizing_function("uniform-l5-0")(H.RobotUtils.build_shape("box-2x2"))
    # lidarSide (RobotUtils.java:244-255) of the normalised cell (1, 0) is E; 5 rays over [-pi/4, pi/4], running sum
    dirs = l5.get(1, 0).get_sensors()[0].sensor.ray_directions
    assert len(dirs) == 5 and dirs[0] == -math.pi / 4 and dirs[2] == pytest.approx(0.0, abs=1e-15) and dirs[4] == pytest.approx(math.pi / 4)
    assert H.Lidar.sample_range_with_rays(1, 1.0, 2.0) == [0.5]      # (end - start) / 2, Lidar.java:89-90
    with pytest.raises(ValueError):
        H.DoubleRange(1.0, 0.0)


def test_tick_times_match_java_accumulation():
    t = H.tick_times(20, 1 / 60)
    g = GOLD["survey_ticks"]
    assert len(t) == g["n"] and t[-1] == g["last"] and t[0] == 1 / 60
    assert np.array_equal(t, oracle.tick_times(20, 1 / 60))
    assert H.abi.load_library().vsr_tick_count(20.0, 1 / 60) == 1200


def test_aggregator_window_lengths():
    # AggregatorSensor.java:56-66 with accumulated double keys: 15/16 and 30/31 samples
    for interval, lens in ((0.25, {15, 16}), (0.5, {30, 31})):
        f = oracle.window_first(20, 1 / 60, interval)
        n = np.arange(len(f)) - f + 1
        assert list(n[:10]) == list(range(1, 11))  # ramp
        assert set(n[60:]) == lens
        # independent restatement in python doubles
        t = H.tick_times(20, 1 / 60)
        ff, ref = 0, []
        for k in range(len(t)):
            while t[ff] < t[k] - interval:
                ff += 1
            ref.append(ff)
        assert np.array_equal(f, ref)


def test_mlp_reference_vectors():
    g = GOLD["reference_mlp_apply"]  # MultiLayerPerceptronTest.java:50-62
    out = oracle.mlp_apply(A.VSR_ACT[g["activation"]], g["neurons"], g["weights"], g["input"])
    assert out.tolist() == g["output"]
    f = GOLD["reference_mlp_flat"]  # :64-94
    assert H.MultiLayerPerceptron.unflat(f["flat"], f["neurons"]) == f["unflat"]
    assert H.MultiLayerPerceptron.flat(f["unflat"], f["neurons"]) == f["flat"]
    assert H.MultiLayerPerceptron.count_weights([32, 21, 21, 16]) == 1507
    assert H.MultiLayerPerceptron.count_weights([32, 16]) == 528
    with pytest.raises(ValueError, match="Wrong number of weights"):
        H.MultiLayerPerceptron("TANH", 3, [2], 1, [0.0] * 5)
    # activation applied to the inputs too (:177): identity weights expose it
    out = oracle.mlp_apply(A.VSR_ACT["TANH"], [1, 1], [0.0, 1.0], [0.5])
    assert out[0] == pytest.approx(math.tanh(math.tanh(0.5)))
    body = helpers.body_of("worm-8x2")
    with pytest.raises(ValueError, match="Wrong dimension of input or output"):
        H.CentralizedSensing(body, H.MultiLayerPerceptron("TANH", 31, [], 16))


def test_compile_batch_layout():
    loco = helpers.locomotion()
    robots = helpers.phase_robots(3) + helpers.phase_robots(2, shape="worm-8x2", seed=5)
    bd = loco.compile(robots)
    d = bd.desc
    assert d.n_robots == 5 and d.n_shapes == 2 and d.controller_kind == A.VSR_CTRL_PHASE_SIN
    assert bd.n_voxels == [10, 16] and bd.n_readings == [20, 32] and bd.n_ticks == 1200
    assert bd.voxel_steps() == (3 * 10 + 2 * 16) * 1200
    assert d.initial_placement == 11.0  # xs[1] + 1, Locomotion.java:50-52
    offs = [d.param_offsets[i] for i in range(6)]
    assert offs == [0, 12, 24, 36, 54, 72]
    assert d.params[0] == 1.0 and d.params[1] == 1.0
    bd1 = loco.compile([helpers.config1_robot()])
    assert bd1.desc.controller_kind == A.VSR_CTRL_TABULATED
    sig = np.ctypeslib.as_array(bd1.desc.params, shape=(1200 * 10,)).reshape(1200, 10)
    t = H.tick_times(20, 1 / 60)
    # voxel order row-major: (0,0),(3,0),(0,1)..; Example.java:47
    assert sig[5, 1] == math.sin(-2 * math.pi * t[5] + math.pi * (3 / 4))
    with pytest.raises(ValueError, match="one controller kind per batch"):
        loco.compile([helpers.config1_robot()] + helpers.phase_robots(1))
    with pytest.raises(ValueError, match="x coordinates must be sorted"):
        H.Locomotion(1.0, [[0, 5, 3], [0, 0, 0]], H.Settings())
    with pytest.raises(ValueError, match="at least 2 points"):
        H.Locomotion(1.0, [[0], [0]], H.Settings())


def test_fp32_protocol_fixture_matches_the_live_oracle():
    """tests/golden/fp32_protocol.npz (the oracle half of the fp32 parity protocol, tests/test_fp32_protocol_gpu.py)
    cannot go stale: a slice of it is recomputed here -- canonical order and one shuffled order."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_fp32_protocol as proto
    from oracle import oracle
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "fp32_protocol.npz"))
    meta = json.loads(str(fix["meta"]))
    assert meta["n"] == proto.N >= 1024 and meta["k"] == proto.K >= 8
    for name, (loco, robots) in proto.configs(8).items():   # the generators are prefix-stable: robot i does not depend on n
        bd = loco.compile(robots[:3])
        r, _ = oracle.run(bd, n_threads=3)
        assert np.allclose(r["last_center_x"] - r["first_center_x"], fix[name + "_canon"][:3], rtol=0, atol=1e-9), name
        s, _ = oracle.run(bd, order=oracle.ORDER_SHUFFLED, seed=meta["shuffle_seed0"] + 2, n_threads=3)
        assert np.allclose(s["last_center_x"] - s["first_center_x"], fix[name + "_shuf"][2][:3], rtol=0, atol=1e-9), name
