"""Observation / VoxelPoly / the per-tick Outcome API and BehaviorUtils (host side),
fed here from the oracle's per-tick probes -- the same arrays the device dumps."""
import importlib
import math

import numpy as np
import pytest

H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers

BU = H.BehaviorUtils


def _square(x0, y0, s, touching=False):
    return H.VoxelPoly([(x0, y0 + s), (x0 + s, y0 + s), (x0 + s, y0), (x0, y0)], 0.0, (0.0, 0.0), touching, 1.0, 0.0)


def test_voxel_poly_geometry():
    p = _square(1.0, 2.0, 3.0)
    assert p.area() == pytest.approx(9.0) and p.center() == pytest.approx((2.5, 3.5))
    b = p.bounding_box()
    assert b.min == (1.0, 2.0) and b.max == (4.0, 5.0)


def test_posture_and_footprint():
    # an L of three voxels (BehaviorUtils.java:177-219: squared box, n x n mask)
    polies = [_square(0, 0, 3, True), _square(3, 0, 3, False), _square(0, 3, 3, False)]
    m = BU.compute_posture(polies, 4)
    assert m.get(0, 0) and m.get(3, 0) and m.get(0, 3) and not m.get(3, 3)
    f = BU.compute_footprint(polies, 4)     # only the first voxel touches: x in [0, 3] of [0, 6];
    assert f.get_mask() == (True, True, True, False) and repr(f) == "___."   # Math.round(1.5) = 2
    avg = BU.compute_average_posture([m, m, BU.compute_posture([_square(0, 0, 3)], 4)])
    assert avg.get(0, 0) and avg.get(0, 3)
    with pytest.raises(ValueError):
        BU.compute_footprint([], 4)


def test_central_element_and_center():
    g = H.Grid.create(3, 2, lambda x, y: None if (x, y) == (1, 0) else (x, y))
    # mean of the non-null cells = (1.0, 0.6): the closest cell is (1, 1)
    assert BU.get_central_element(g) == (1, 1)
    assert BU.center([_square(0, 0, 2), _square(2, 0, 2)]) == pytest.approx((2.0, 1.0))
    with pytest.raises(ValueError):
        BU.get_central_element(H.Grid.create(2, 2, None))


def test_spectrum_follows_the_reference_formulas():
    # 1 Hz sine, 60 samples/s: padded to 1024, |FFT| of the first 513 bins, f_i = 1/dT/2 * i / 513
    n, dt = 600, 1 / 60
    sig = {k * dt: math.sin(2 * math.pi * k * dt) for k in range(n)}
    sp = BU.compute_spectrum(sig)
    assert len(sp) == 513
    peak = max(sp, key=sp.get)
    assert peak == pytest.approx(30.0 * 17 / 513)
    q = BU.compute_quantized_spectrum(sig, 0.0, 5.0, 10)
    assert list(q)[0] == (0.0, 0.5) and max(q, key=q.get) == (0.5, 1.0)


def test_gaits_from_a_periodic_footprint_sequence():
    a, b = H.Footprint([True, False]), H.Footprint([False, True])
    fps = {0.5 * k: (a if k % 2 == 0 else b) for k in range(12)}
    gaits = BU.compute_gaits(fps, 2, 4, 0.5)
    assert gaits and all(g.get_mode_interval() == gaits[0].get_mode_interval() for g in gaits)
    g = max(gaits, key=lambda g: g.get_duration())
    assert len(g.get_footprints()) >= 2 and 0.0 < g.get_avg_touch_area() <= 1.0 and g.get_purity() > 0


def test_observations_from_per_tick_dumps_match_the_result_record():
    loco = helpers.locomotion(3.0)
    robot = helpers.config1_robot()
    bd = loco.compile([robot])
    res, pr = oracle.run(bd, trajectory=True, signals=True)
    nv = bd.n_voxels[0]
    trace = np.stack([pr["touch"][0][:, :nv], pr["signals"][0][:, :nv]], axis=-1)
    mat = H.abi.default_material()
    obs = H.build_observations(robot.get_voxels(), pr["trajectory"][0][:, :4 * nv], trace, H.tick_times(3.0, 1 / 60),
                               loco.ground_profile, mat.side_length, mat.side_length * mat.mass_side_length_ratio)
    out = H.Outcome(observations=obs)
    assert len(out.get_observations()) == 181
    r = res[0]
    assert out.get_distance() == pytest.approx(r["last_center_x"] - r["first_center_x"], abs=1e-9)
    assert out.get_time() == pytest.approx(r["t_last"] - r["t_first"], abs=1e-12)
    assert out.get_control_energy() == pytest.approx(r["control_energy_last"] - r["control_energy_first"], rel=1e-9)
    assert out.get_area_ratio_energy() == pytest.approx(r["area_ratio_energy_last"] - r["area_ratio_energy_first"], rel=1e-9)
    first = next(iter(obs.values()))
    assert first.terrain_height == 5.0                       # flat: Ground.yAt under the robot
    assert sum(v is not None for v in first.voxel_polies.values()) == nv
    # a biped standing/hopping on flat ground: something touches at some tick, the posture is solid
    assert any(v.is_touching_ground() for o in obs.values() for v in o.voxel_polies.values() if v is not None)
    post = out.get_average_posture(8)
    assert post.count(lambda b: b) > 16
    spec = out.get_center_y_position_spectrum(0.0, 5.0, 10)
    assert len(spec) == 10 and all(v >= 0 for v in spec.values())
    assert len(out.get_footprints_spectra(4, 0.0, 5.0, 5)) == 4
    half = out.sub_outcome(1.0, 2.0)
    assert 59 <= len(half.get_observations()) <= 61 and half.get_time() < 1.0
    with pytest.raises(NotImplementedError):
        H.Outcome(res[0]).get_observations()
