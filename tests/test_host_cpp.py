"""The C++ host mirror (2dhmsr_b200/host/hmsrobots.hpp) against the reference's vectors, the Python
mirror (two independent implementations of RobotUtils / createTerrain / java.util.Random) and,
on a GPU, against the Python path through the same C ABI."""
import importlib
import json
import math
import os
import subprocess

import numpy as np
import pytest

H = importlib.import_module("2dhmsr_b200")
A = H.abi
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "2dhmsr_b200", "host", "host_check")
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def _run(mode):
    out = subprocess.run([EXE, mode], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout)


def test_cpp_host_mirror_selfcheck():
    j = _run("selfcheck")
    g = GOLD["survey_java_random"]
    assert j["seed0_nextDouble"] == g["seed0_nextDouble"] and j["seed0_nextInt"] == g["seed0_nextInt"]
    assert j["seed42_nextInt"] == g["seed42_nextInt"]
    assert j["seed0_nextGaussian"] == pytest.approx(g["seed0_nextGaussian"], abs=1e-15)
    # terrain DSL: identical to the Python mirror (independent implementation) and to the SURVEY points
    for name, kx, ky in (("hilly-1-10-0", "hilly_xs", "hilly_ys"), ("steppy-2-8-3", "steppy_xs", "steppy_ys")):
        t = H.Locomotion.create_terrain(name)
        assert np.allclose(j[kx], t[0], rtol=0, atol=1e-12) and np.allclose(j[ky], t[1], rtol=0, atol=1e-12)
    assert len(j["hilly_xs"]) == GOLD["survey_hilly_1_10_0"]["n_points"]
    assert np.allclose(j["flatWithStart_ys"], H.Locomotion.create_terrain("flatWithStart-2")[1], atol=1e-12)
    assert np.allclose(j["uphill_ys"], H.Locomotion.create_terrain("uphill-10")[1], atol=1e-12)
    for name, m in j["shapes"].items():
        ref = "".join("1" if v else "0" for v in H.RobotUtils.build_shape(name).values())
        assert m == ref, name
    # MultiLayerPerceptronTest.java:50-95
    assert j["mlp_apply"] == GOLD["reference_mlp_apply"]["output"]
    assert j["mlp_unflat_layer0_neuron1"] == GOLD["reference_mlp_flat"]["unflat"][0][1]
    assert j["mlp_flat"] == GOLD["reference_mlp_flat"]["flat"] and j["count_weights"] == 1507
    assert j["n_ticks"] == 1200 and j["last_tick"] == GOLD["survey_ticks"]["last"]
    # descriptors equal the Python mirror's
    bd = helpers.locomotion().compile([helpers.config1_robot()])
    assert j["cfg1_kind"] == bd.desc.controller_kind and j["cfg1_n_params"] == 12000
    sig = np.ctypeslib.as_array(bd.desc.params, shape=(12000,))
    assert j["cfg1_param_5_1"] == sig[51] and j["cfg1_initial_placement"] == 11.0
    enc = []
    for i in range(bd.desc.shapes[0].sensor_offsets[10]):
        s = bd.desc.shapes[0].sensors[i]
        enc += [s.base, s.axes, s.rotated, s.aggregator, s.soft_norm, s.max_velocity_norm, s.interval]
    assert j["cfg1_sensors"] == enc
    assert j["phase_n_shapes"] == 1 and len(j["phase_params"]) == 3 * 12
    assert j["phase_params"][:4] == [1.0, 1.0, 0.1, 0.1 + 0.7 * 3]  # voxel order (0,0), (3,0), ...
    assert j["worm_inputs"] == 64
    # DistributedSensing: biped-4x3 x (2 readings + 4 signals -> [2] -> 1 + 4): 10 * (2*7 + 5*3) weights
    assert j["dist_kind"] == A.VSR_CTRL_DISTRIBUTED and j["dist_signals"] == 1 and j["dist_n_params"] == 290
    assert j["dist_param_0"] == H.JavaRandom(7).next_double() * 2.0 - 1.0
    assert j["dist_nd_kind"] == A.VSR_CTRL_DISTRIBUTED_ND and j["dist_nd_n_params"] == 10 * (2 * 7 + 2 * 3)
    assert j["spined_sensor_counts"] == [3, 3, 2, 2, 2, 2, 3, 3, 3, 4]
    # uniform-r+px+py+cpg+m at (3, 1) of a 4x2 worm: (base, normalisation, parameter) per sensor
    assert j["rich_sensors_3_1"] == [A.VSR_SENSOR_ANGLE, 1, 0, A.VSR_SENSOR_CONSTANT, 0, 1.0, A.VSR_SENSOR_CONSTANT, 0, 1.0,
                                     A.VSR_SENSOR_TIME_SIN, 2, -1.0, A.VSR_SENSOR_MALFUNCTION, 0, 0]
    # error behaviour: IllegalArgumentException -> std::invalid_argument, same messages
    assert j["err_terrain"] == "invalid_argument:Unknown terrain name: moon"
    assert j["err_shape"] == "invalid_argument:Unknown body name: snake-3"
    assert j["err_sensors"].startswith("invalid_argument:Unknown sensorizing function name")
    # uniformAll (RobotUtils.java:177-189): 15 sensors in TreeMap order, 21 readings per voxel
    allv = helpers.body_of("box-2x2", "uniformAll-0").get(1, 1)
    assert j["all_bases_1_1"] == [s.encode().base for s in allv.get_sensors()] and len(j["all_bases_1_1"]) == 15
    assert j["all_readings"] == sum(len(s.domains()) for s in allv.get_sensors()) == 21
    # spinedTouchSighted (:150-165): spinedTouch + a 5-ray lidar on the last column
    sighted = helpers.body_of("biped-4x3", "spinedTouchSighted-t-f-0")
    assert j["sighted_sensor_counts"] == [len(v.get_sensors()) for v in sighted.values() if v is not None]
    enc = helpers.body_of("worm-3x2", "uniform-l5-0").get(2, 0).get_sensors()[0].encode()
    assert j["lidar_2_0"] == [A.VSR_SENSOR_LIDAR, 5, A.VSR_NORM_LINEAR, 10.0] + list(enc.lidar_dirs)[:5]
    assert j["err_weights"] == "invalid_argument:Wrong number of weights: 11 expected, 5 found"
    assert j["err_dims"].startswith("invalid_argument:Wrong dimension of input or output")
    assert j["err_ground"] == "invalid_argument:x coordinates must be sorted"
    assert j["err_range"].startswith("invalid_argument:Max has to be lower")


@pytest.mark.gpu
def test_cpp_locomotion_apply_equals_python_path():
    j = _run("run")
    default = lambda: H.Locomotion(20.0, H.Locomotion.create_terrain("flat"), H.Settings())   # as hmsrobots::Settings()
    o = default().apply(helpers.config1_robot())
    assert j["cfg1"]["distance"] == o.get_distance() and j["cfg1"]["velocity"] == o.get_velocity()
    assert j["cfg1"]["time"] == pytest.approx(20 - 1 / 60, abs=1e-9)
    body = helpers.body_of("biped-4x3")
    robots = [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(4, 3, lambda x, y, b=b: b + 0.7 * x + 0.3 * y)), body)
              for b in (0.1, 0.2, 0.3)]
    outs = default().apply(robots)
    assert j["phase_distance"] == [o.get_distance() for o in outs]

    def dist_robot(seed, directional):
        body = helpers.body_of("biped-4x3")
        ds = (H.DistributedSensing if directional else H.DistributedSensingNonDirectional)(body, 1)
        rnd = H.JavaRandom(seed)
        for (x, y), v in body:
            if v is None:
                continue
            nin, nout = ds.n_of_inputs(x, y), ds.n_of_outputs(x, y)
            nw = H.MultiLayerPerceptron.count_weights([nin, 2, nout])
            ds.get_functions().set(x, y, H.MultiLayerPerceptron("TANH", nin, [2], nout, [rnd.next_double() * 2.0 - 1.0 for _ in range(nw)]))
        return H.Robot(ds, body)

    l5 = H.Locomotion(5.0, H.Locomotion.create_terrain("flat"), H.Settings())
    want = [o.get_distance() for o in l5.apply([dist_robot(7, True), dist_robot(8, True)])] + \
           [l5.apply(dist_robot(7, False)).get_distance()]
    assert j["distributed_distance"] == want
    # the per-tick Observation map through the C++ mirror against the Python one
    o = H.Locomotion(2.0, H.Locomotion.create_terrain("flat"), H.Settings()).apply(robots[0], observations=True)
    obs = o.get_observations()
    polies = lambda ob: [v for v in ob.voxel_polies.values() if v is not None]
    last = list(obs.values())[-1]
    assert j["obs"]["n"] == len(obs) == 121
    assert j["obs"]["distance"] == pytest.approx(o.get_distance(), abs=1e-9)
    assert j["obs"]["record_distance"] == pytest.approx(j["obs"]["distance"], abs=1e-4)   # fp32 record vs double host sum
    assert j["obs"]["touches"] == sum(p.is_touching_ground() for ob in obs.values() for p in polies(ob))
    assert j["obs"]["last_control_energy"] == pytest.approx(polies(last)[0].get_control_energy(), rel=1e-12)
    assert j["obs"]["terrain_height"] == last.terrain_height == 5.0
    assert j["obs"]["footprint"] == repr(H.BehaviorUtils.compute_footprint(polies(last), 4))
    assert j["obs"]["sub"] == len(o.sub_outcome(0.5, 1.0).get_observations())
