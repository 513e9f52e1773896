"""The oracle against closed forms, its own regression pins and the reference's shipped frames."""
import importlib
import math
import json
import os

import numpy as np
import pytest

H = importlib.import_module("2dhmsr_b200")
from oracle import oracle
import helpers

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
FRAME_TICKS = (240, 253, 266, 279, 292)  # FramesImageBuilder.java:85-92 on the accumulated tick times


def test_frame_ticks_follow_the_reference_listener():
    t = H.tick_times(20, 1 / 60)
    out, last = [], -1e9
    for k, tt in enumerate(t):
        if tt < 4 or tt >= 5 or tt - last < 0.2:
            continue
        last = tt
        out.append(k)
    assert tuple(out) == FRAME_TICKS
    # InfoDrawer prints %4.1f: the labels on assets/frames.png
    assert ["%.1f" % t[k] for k in out] == [s.replace(",", ".") for s in GOLD["reference_frames_png"]["labels"]]


def test_free_fall_matches_closed_form():
    # before the first contact all springs are at rest: v <- (v + g dt)(1 - 0.1 dt), y <- y + v dt
    bd = helpers.locomotion(1.0).compile(helpers.phase_robots(1, amp=0.0))
    _, pr = oracle.run(bd, trajectory=True, stats=True)
    tr = pr["trajectory"][0]
    dt, v, y = 1 / 60, 0.0, None
    y0 = tr[0, :, 1] - (0 + (-9.8 * dt) * (1 - 0.1 * dt)) * dt
    y, v = y0.copy(), 0.0
    first_contact = int(np.argmax(pr["n_contacts"][0] > 0))
    assert first_contact > 20
    for k in range(first_contact):
        v = (v + -9.8 * dt) * (1 - 0.1 * dt)
        y = y + v * dt
        assert np.abs(tr[k, :, 1] - y).max() < 1e-12
        assert np.abs(tr[k, :, 0] - tr[0, :, 0]).max() < 1e-12 and np.abs(tr[k, :, 2]).max() < 1e-12
    # placement: lowest voxel 1.0 above the ground (Locomotion.java:183-187); bbox.min.x = 11 (:180-181)
    assert y0.min() - 0.525 == pytest.approx(5.0 + 1.0, abs=1e-12)
    assert tr[0, :, 0].min() - 0.525 == pytest.approx(11.0, abs=1e-12)


def test_single_voxel_settles():
    bd = helpers.locomotion(5.0).compile(helpers.phase_robots(1, shape="box-1x1", sensors="uniform-a+t-0", amp=0.0))
    res, pr = oracle.run(bd, trajectory=True, readings=True, stats=True)
    tr = pr["trajectory"][0][:, :4]
    assert np.abs(tr[-1, :, 3:]).max() < 1e-3            # at rest (residual creep of the soft springs)
    assert tr[-1, :, 1].min() - 0.525 == pytest.approx(5.0 - 0.005, abs=2e-4)  # rests at -linearTolerance
    area, touch = pr["readings"][0, -1, :2]
    assert area == pytest.approx(0.5, abs=5e-3)           # SoftNormalization(area ratio 1) = 0.5
    assert touch == 1.0 and pr["readings"][0, 0, 1] == 0.0
    assert res[0]["status"] == 0 and pr["n_contacts"][0, -1] == 2


def test_symmetric_robot_does_not_drift():
    bd = helpers.locomotion(5.0).compile(helpers.phase_robots(1, shape="box-2x2", amp=0.0))
    res, _ = oracle.run(bd)
    # the Gauss-Seidel sweep is not mirror-symmetric, so a few millimetres of creep are expected
    assert abs(res[0]["last_center_x"] - res[0]["first_center_x"]) < 2e-2


def test_outcome_fields():
    bd = helpers.locomotion(2.0).compile(helpers.phase_robots(2, seed=3))
    res, pr = oracle.run(bd, signals=True)
    o = H.Outcome(res[0])
    t = H.tick_times(2.0, 1 / 60)  # 121 ticks: the accumulated t reaches 1.9999999999999978 < 2 at tick 120
    assert len(t) == 121 and res[0]["n_ticks"] == 121
    assert o.get_time() == t[-1] - t[0]  # Outcome.java:194-196 (not finalT)
    assert o.get_velocity() == pytest.approx(o.get_distance() / o.get_time())
    # controlEnergy at tick k adds the force applied at tick k-1 (Voxel.java:198-203)
    f = pr["signals"][0]
    assert res[0]["control_energy_first"] == 0.0
    assert res[0]["control_energy_last"] == pytest.approx((f[:-1] ** 2).sum(), rel=1e-12)
    assert np.abs(f).max() <= 1.0


def test_oracle_regression_pins():
    loco = helpers.locomotion()
    res, pr = oracle.run(loco.compile([helpers.config1_robot()]), trajectory=True)
    g = GOLD["oracle_config1"]
    for k, v in g["result"].items():
        assert float(res[0][k]) == pytest.approx(v, rel=1e-9, abs=1e-9), k
    cx, cy = helpers.poly_centres(pr["trajectory"][0])
    assert np.allclose([[cx[k], cy[k]] for k in FRAME_TICKS], g["centre_at_frames"], atol=1e-7)
    res, _ = oracle.run(helpers.locomotion(5.0).compile(helpers.phase_robots(3, seed=7)))
    assert np.allclose(res["last_center_x"] - res["first_center_x"], GOLD["oracle_phase3"]["distance"], atol=1e-8)
    res, _ = oracle.run(helpers.locomotion(3.0).compile(helpers.mlp_robots(2, hidden=(5,), seed=3)))
    assert np.allclose(res["last_center_x"] - res["first_center_x"], GOLD["oracle_mlp2"]["distance"], atol=1e-8)


def test_threads_and_ranges_agree():
    bd = helpers.locomotion(2.0).compile(helpers.phase_robots(6, seed=11))
    a, _ = oracle.run(bd, n_threads=1)
    b, _ = oracle.run(bd, n_threads=4)
    c, _ = oracle.run(bd, robot_begin=2, robot_end=5)
    assert a.tobytes() == b.tobytes() and a[2:5].tobytes() == c.tobytes()


def test_order_sensitivity_is_bounded_early():
    # Gauss-Seidel order differs between dyn4j (DFS island order) and the GPU colouring: over the
    # first second the three orders stay within millimetres of each other.
    bd = helpers.locomotion(1.0).compile(helpers.phase_robots(4, seed=2))
    tr = [oracle.run(bd, order=o, seed=5, trajectory=True)[1]["trajectory"] for o in (0, 1, 2)]
    # measured: ~0.03 after 20 ticks, ~0.5 after 60 ticks -- 10 sequential-impulse iterations do not
    # converge the stiff actuated springs, so the sweep order is a first-order effect (this is why
    # north_star only asks for short-horizon / distribution-level agreement with dyn4j's own order)
    for other in (1, 2):
        assert np.abs(tr[0][:, :20, :, :2] - tr[other][:, :20, :, :2]).max() < 0.1
        assert np.abs(tr[0][..., :2] - tr[other][..., :2]).max() < 2.0


def test_spring_mass_modes_and_validation():
    r = helpers.phase_robots(1, amp=0.0)
    a, _ = oracle.run(helpers.locomotion(3.0, spring_mass_mode=0).compile(r))
    b, _ = oracle.run(helpers.locomotion(3.0, spring_mass_mode=1).compile(r))
    assert a[0]["last_center_y"] != b[0]["last_center_y"]  # the switch is live
    m2 = H.abi.default_material()
    m2.area_ratio_passive_min = 0.2
    with pytest.raises(ValueError, match="areaRatioPassiveRange"):
        H.Voxel([], m2)  # Voxel.java:132-140


def _frames_residual(**kw):
    sc = kw.pop("self_collision", 1)
    sm = kw.pop("spring_mass_mode", 0)
    bd = helpers.locomotion(5.0, spring_mass_mode=sm, self_collision=sc).compile([helpers.config1_robot()])
    _, pr = oracle.run(bd, trajectory=True, self_collision=bool(sc), **kw)
    cx, cy = helpers.poly_centres(pr["trajectory"][0])
    g = np.asarray(GOLD["reference_frames_png"]["centres"])
    return np.abs(np.asarray([[cx[k], cy[k]] for k in FRAME_TICKS]) - g).max()


@pytest.mark.xfail(strict=True, reason="dyn4j boundary is parity-unpinned: no combination of the oracle's switches (spring "
                   "mass convention x intra-robot contacts x CCD x constraint order, incl. dyn4j's island DFS order) "
                   "reproduces the five robot-centre positions printed on the reference's assets/frames.png; the "
                   "residual table is in DESIGN.md section 2")
def test_reference_frames_png_centres():
    best = min(_frames_residual(spring_mass_mode=sm, self_collision=sc, ccd=bool(ccd), order=order)
               for sm in (0, 1) for sc in (1, 0) for ccd in (0, 1) for order in (oracle.ORDER_CANONICAL, oracle.ORDER_DFS))
    assert best <= 0.3


def test_frames_png_residuals_are_the_tabulated_ones():
    # DESIGN.md section 2 quotes these (scripts/frames_ensemble.py prints the whole table): the trajectory is chaotic
    # after ~2 s, so the residual of ONE robot at t = 4 s moves by units when nothing but the solver order changes
    assert _frames_residual() == pytest.approx(5.83, abs=0.02)                                  # the default switches
    assert _frames_residual(order=oracle.ORDER_DFS) == pytest.approx(4.03, abs=0.02)            # dyn4j's island order
    assert _frames_residual(self_collision=0, ccd=True) == pytest.approx(2.87, abs=0.02)        # the best of the 24 rows


def test_dyn4j_contact_labels_switch():
    # the third dyn4j item behind a switch: contact ids labelled as dyn4j does (the wrap-around edge
    # is 0 or n by a floating-point tie) instead of one label per edge.  It only changes which impulses are warm-started:
    # live, deterministic, and it does not close the frames.png gap either.
    bd = helpers.locomotion(3.0).compile(helpers.phase_robots(4, seed=2))
    a, _ = oracle.run(bd)
    b2, _ = oracle.run(bd, dyn4j_labels=True)
    c, _ = oracle.run(bd, dyn4j_labels=True)
    assert b2.tobytes() == c.tobytes() and a.tobytes() != b2.tobytes()
    assert np.abs(a["last_center_x"] - b2["last_center_x"]).max() < 3.0
    assert _frames_residual(dyn4j_labels=True) > 2.0 and _frames_residual(dyn4j_labels=True, order=oracle.ORDER_DFS, ccd=True) > 2.0


def test_continuous_detection_setting_is_the_ccd_switch():
    # vsr_settings.continuous_detection (what the kernel reads) and vsr_oracle_opts.ccd are the same pass
    robots = helpers.phase_robots(3, seed=2)
    a, _ = oracle.run(helpers.locomotion(2.0).compile(robots), ccd=True)
    b2, _ = oracle.run(helpers.locomotion(2.0, continuous_detection=1).compile(robots))
    c, _ = oracle.run(helpers.locomotion(2.0).compile(robots))
    assert a.tobytes() == b2.tobytes() and a.tobytes() != c.tobytes()


def test_first_touch_of_a_dropped_voxel_under_continuous_detection():
    # A known answer for the time-of-impact pass: a passive voxel dropped on flat ground (y = 5).  Without the pass the
    # bottom masses are first seen 0.0126 INSIDE the ground (they fall 0.072 per tick from 0.059 above it).  With it
    # their first touch is linearTolerance deep: the conservative advancement of a mass that does not rotate lands
    # exactly on the surface, which the "touching = not separated" rule reads as an overshoot -- the
    # mass is held at its pose for that tick (time of impact 0, as dyn4j does whenever Gjk.distance reports the landing
    # as not separated) -- and the tick after it is put linearTolerance = 0.005 into the ground by TimeOfImpactSolver.
    rob = helpers.phase_robots(1, shape="box-1x1", amp=0.0)
    half, out = 0.525, {}
    for ccd in (0, 1):
        _, pr = oracle.run(helpers.locomotion(1.5, continuous_detection=ccd).compile(rob), trajectory=True, stats=True)
        tr = pr["trajectory"][0]
        bottom = (tr[:, :4, 1] - half * (np.abs(np.cos(tr[:, :4, 2])) + np.abs(np.sin(tr[:, :4, 2])))).min(axis=1) - 5.0
        out[ccd] = (bottom, int(np.argmax(pr["n_contacts"][0] > 0)))
    b0, f0 = out[0]
    b1, f1 = out[1]
    assert f1 == f0 + 1 and np.array_equal(b0[:f0], b1[:f0])               # identical until the tick of the first touch
    assert b0[f0 - 1] > 0.05 and b0[f0] == pytest.approx(-0.0126, abs=2e-4)
    assert b1[f0] == b1[f0 - 1]                                            # held for one tick ...
    assert b1[f1] == pytest.approx(-0.005, abs=2e-5)                       # ... then linearTolerance deep


def test_the_two_readings_of_a_touching_pair_give_the_same_population():
    # dyn4j's Gjk.distance may report an exactly touching pair either way.  The oracle's default --
    # which the kernel shares -- is "not separated"; vsr_oracle_opts.touching_separated is the other reading (the
    # advance stands, the mass is pushed in the same tick).  They differ at the first touch of a dropped voxel (one tick
    # earlier) and nowhere that matters: 1024 robots of config 2 over 20 s give 11.741 and 11.733 (reorder null +-0.52),
    # rank correlation 0.999 -- here on the first 256 of them.
    rob = helpers.phase_robots(1, shape="box-1x1", amp=0.0)
    _, pr = oracle.run(helpers.locomotion(1.5, continuous_detection=1).compile(rob), trajectory=True, stats=True,
                       touching_separated=True)
    tr = pr["trajectory"][0]
    bottom = (tr[:, :4, 1] - 0.525 * (np.abs(np.cos(tr[:, :4, 2])) + np.abs(np.sin(tr[:, :4, 2])))).min(axis=1) - 5.0
    f = int(np.argmax(pr["n_contacts"][0] > 0))
    assert bottom[f - 1] > 0.05 and bottom[f] == pytest.approx(-0.005, abs=2e-5)   # linearTolerance deep in the tick it lands
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_fp32_protocol as proto
    from scipy.stats import spearmanr
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "fp32_protocol.npz"))
    n = 256
    loco, robots = proto.configs(n)["cfg2"]
    r, _ = oracle.run(loco.compile(robots), n_threads=os.cpu_count(), touching_separated=True)
    d, canon, shuf = r["last_center_x"] - r["first_center_x"], fix["cfg2_canon"][:n], fix["cfg2_shuf"][:, :n]
    null_mean = max(abs(shuf[k].mean() - canon.mean()) for k in range(shuf.shape[0]))
    assert np.all(r["status"] == 0)
    assert spearmanr(d, canon)[0] > 0.98 and abs(d.mean() - canon.mean()) < 0.25 * null_mean


def test_island_dfs_order_and_ccd_switches():
    # dyn4j's own constraint order (island discovery, depth first) and its time-of-impact pass sit
    # behind single switches of the oracle; both are deterministic, both change a trajectory only after the first
    # ground approach, and neither moves it by more than the solver-order sensitivity early on
    bd = helpers.locomotion(1.5).compile(helpers.phase_robots(4, seed=2))
    _, base = oracle.run(bd, trajectory=True, stats=True)
    _, dfs = oracle.run(bd, trajectory=True, order=oracle.ORDER_DFS)
    _, dfs2 = oracle.run(bd, trajectory=True, order=oracle.ORDER_DFS)
    _, ccd = oracle.run(bd, trajectory=True, ccd=True)
    assert dfs["trajectory"].tobytes() == dfs2["trajectory"].tobytes()
    first = int(np.argmax(base["n_contacts"].max(0) > 0))
    d_dfs = np.abs(dfs["trajectory"][..., :2] - base["trajectory"][..., :2]).max((0, 2, 3))
    d_ccd = np.abs(ccd["trajectory"][..., :2] - base["trajectory"][..., :2]).max((0, 2, 3))
    assert 0 < d_dfs[20] < 0.1 and d_dfs[60] < 1.5                      # order matters from the first actuated tick on
    assert d_ccd[:10].max() == 0.0 and d_ccd.max() > 0                  # CCD: nothing before a mass nears the ground
    assert d_ccd[min(first + 30, len(d_ccd) - 1)] < 1.5


def test_dfs_order_stays_inside_the_reorder_null_and_ccd_does_not():
    # What two dyn4j features do to the 20 s outcome of BASELINE config 2, measured with the statistics of the fp32
    # protocol (tests/test_fp32_protocol_gpu.py) on its fixture.  The island DFS order (C.6, oracle only) moves the
    # population no more than a seeded reshuffle of the order does.  Continuous detection (C.7;
    # vsr_settings.continuous_detection, on by default as in dyn4j, in the kernel too) does NOT stay inside that null:
    # switched off, the mean displacement of 1024 robots is 8 % higher (11.74 -> 12.70, 2.5 x the largest reshuffle),
    # while the ranking of the robots is kept as well as a reshuffle keeps it.  (configs 3 and 4: 3.21 -> 3.67 and
    # 8.13 -> 8.75)  This is why the pass is not a "documented omission"  here.
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_fp32_protocol as proto
    from scipy.stats import spearmanr
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "fp32_protocol.npz"))
    for cfg in ("cfg2", "cfg3", "cfg4"):
        canon, nopass, shuf = fix[cfg + "_canon"], fix[cfg + "_canon_nopass"], fix[cfg + "_shuf"]
        null_mean = max(abs(shuf[k].mean() - canon.mean()) for k in range(shuf.shape[0]))
        null_rho = min(spearmanr(shuf[k], canon)[0] for k in range(shuf.shape[0]))
        shift = nopass.mean() - canon.mean()
        assert 1.2 * null_mean < shift < 0.2 * canon.mean(), (cfg, shift, null_mean)
        assert spearmanr(nopass, canon)[0] >= null_rho - 0.05, cfg
    n = 256
    loco, robots = proto.configs(n)["cfg2"]
    canon, shuf = fix["cfg2_canon"][:n], fix["cfg2_shuf"][:, :n]
    null_mean = max(abs(shuf[k].mean() - canon.mean()) for k in range(shuf.shape[0]))
    null_rho = min(spearmanr(shuf[k], canon)[0] for k in range(shuf.shape[0]))
    r, _ = oracle.run(loco.compile(robots), n_threads=os.cpu_count(), order=oracle.ORDER_DFS)
    d = r["last_center_x"] - r["first_center_x"]
    assert np.all(r["status"] == 0)
    assert spearmanr(d, canon)[0] >= null_rho - 0.05
    assert abs(d.mean() - canon.mean()) <= max(0.05 * abs(canon.mean()), 1.5 * null_mean)


def test_intra_robot_contacts_follow_the_setting():
    # Voxel.java:178-185,474: every pair of masses of a robot may collide unless welded.  At rest the
    # corner neighbours only touch (no constraint: a contact needs a penetration above the tie margin);
    # once the voxels deform they overlap.  self_collision=0 keeps the ground contacts only.
    robots = helpers.phase_robots(4, seed=5)
    on, pon = oracle.run(helpers.locomotion(1.0, self_collision=1).compile(robots), stats=True)
    off, poff = oracle.run(helpers.locomotion(1.0, self_collision=0).compile(robots), stats=True)
    assert pon["n_contacts"][:, 0].max() == 0
    assert poff["n_contacts"][:, :15].max() == 0          # free fall: nothing touches the ground yet
    assert pon["n_contacts"][:, 1:15].max() > 0           # ... but corner neighbours already overlap
    assert np.all(on["status"] == 0) and np.all(off["status"] == 0)
    assert np.abs(on["last_center_y"] - off["last_center_y"]).max() > 1e-9
    # a single voxel has no non-welded neighbours close enough: identical results
    one = helpers.phase_robots(2, shape="box-1x1", seed=5)
    a, _ = oracle.run(helpers.locomotion(2.0, self_collision=1).compile(one))
    b, _ = oracle.run(helpers.locomotion(2.0, self_collision=0).compile(one))
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("directional", [True, False])
def test_distributed_sensing_matches_the_host_restatement(directional):
    # SURVEY 8(f) rank 2.  The oracle's per-tick readings fed through the Python DistributedSensing
    # (a line-by-line restatement of DistributedSensing.java:150-207 over MultiLayerPerceptron.apply)
    # must give the forces the oracle applied.
    robots = helpers.distributed_robots(2, signals=2, hidden=(3,), seed=9, directional=directional)
    bd = helpers.locomotion(1.0).compile(robots)
    res, pr = oracle.run(bd, signals=True, readings=True)
    assert np.all(res["status"] == 0)
    body = robots[0].get_voxels()
    cells = [(x, y) for (x, y), v in body if v is not None]
    for i, r in enumerate(robots):
        ctrl = r.get_controller()
        ctrl.reset()
        for k in range(0, 60):
            rd = H.Grid.create(body.w, body.h, None)
            for j, (x, y) in enumerate(cells):
                rd.set(x, y, list(pr["readings"][i, k, 2 * j:2 * j + 2]))   # uniform-ax+t-0: 2 readings per voxel
            out = ctrl.compute_control_signals(0.0, rd)
            want = [max(-1.0, min(1.0, out.get(x, y))) for (x, y) in cells]   # Voxel.applyForce clamps
            assert pr["signals"][i, k, :len(cells)] == pytest.approx(want, abs=1e-12), k
    # signals matter: the same weights with the neighbours' messages cut (signals = 0 is another network)
    assert np.abs(pr["signals"][0, 30] - pr["signals"][0, 1]).max() > 1e-6
    with pytest.raises(NotImplementedError):
        broken = helpers.distributed_robots(1)[0]
        broken.get_controller().get_functions().set(0, 0, None)
        helpers.locomotion(1.0).compile([broken])


def test_cheap_sensors_read_what_the_reference_defines():
    # Angle, Constant (px, py), Malfunction, TimeFunction ("cpg") and the linear
    # Normalization wrapper (RobotUtils.java:41,54-57)
    body = helpers.body_of("worm-4x2", "uniform-r+px+py+cpg+m-0")
    robots = [H.Robot(H.PhaseSin(1.0, 0.5, H.Grid.create(4, 2, lambda x, y: 0.3 * x)), body)]
    loco = helpers.locomotion(1.0)
    bd = loco.compile(robots)
    res, pr = oracle.run(bd, readings=True, trajectory=True)
    rd = pr["readings"][0].reshape(60, 8, 5)          # [tick][voxel][r, px, py, cpg, m]
    t = H.tick_times(1.0, 1 / 60)
    cells = [(x, y) for (x, y), v in body if v is not None]
    for j, (x, y) in enumerate(cells):
        assert np.all(rd[:, j, 1] == x / 3.0) and np.all(rd[:, j, 2] == y / 1.0)    # Constant(x/(w-1)), Constant(y/(h-1))
        assert np.all(rd[:, j, 4] == 0.0)                                            # never broken
        assert rd[:, j, 3] == pytest.approx((np.sin(2 * np.pi * -1 * t) + 1) / 2, abs=1e-12)   # Normalization of [-1, 1]
    # "r" = SoftNormalization(Angle): angle 0 -> 0.5 at the first tick, then the actuation tilts the voxels a little
    assert rd[0, :, 0] == pytest.approx(0.5, abs=1e-9) and rd[:10, :, 0] == pytest.approx(0.5, abs=5e-3)
    tr = pr["trajectory"][0][:, :32].reshape(60, 8, 4, 6)
    ang = (np.arctan2(tr[:, :, 1, 1] - tr[:, :, 0, 1], tr[:, :, 1, 0] - tr[:, :, 0, 0]) +
           np.arctan2(tr[:, :, 2, 1] - tr[:, :, 3, 1], tr[:, :, 2, 0] - tr[:, :, 3, 0])) / 2
    want = np.tanh(((np.clip(ang, -np.pi, np.pi) + np.pi) / (2 * np.pi) * 2 - 1) * 2) / 2 + 0.5
    assert rd[:, :, 0] == pytest.approx(want, abs=1e-12)
    # AppliedForce reads the force of the previous control step
    body2 = H.Grid.create(2, 1, lambda x, y: H.Voxel([H.AppliedForce()]))
    r2 = [H.Robot(H.PhaseSin(1.0, 0.7, H.Grid.create(2, 1, 0.0)), body2)]
    _, p2 = oracle.run(loco.compile(r2), readings=True, signals=True)
    assert p2["readings"][0, 0, :2] == pytest.approx(0.0) and \
        p2["readings"][0, 1:, :2] == pytest.approx(p2["signals"][0, :-1, :2], abs=1e-15)
    assert [len(v.get_sensors()) for v in helpers.body_of("biped-4x3", "spinedTouch-t-t-0").values() if v] == \
        [3, 3, 2, 2, 2, 2, 3, 3, 3, 4]
    with pytest.raises(NotImplementedError):
        helpers.body_of("biped-4x3", "spinedTouchSighted-t-t-0") and H.Lidar(10.0)     # a lidar needs 1..8 rays


def test_noisy_sensors_draw_from_java_util_random():
    # Noisy.java:49-62: every sensor instance owns a Random(0); reading i of a sensor of dimension d gets
    # gaussian number tick * d + i of that stream, scaled by sigma * extent of the (outer) domain.
    sigma = 0.05
    clean = helpers.body_of("worm-3x1", "uniform-t+vxy+px-0")
    noisy = helpers.body_of("worm-3x1", f"uniform-t+vxy+px-{sigma}")
    ctrl = lambda b: H.PhaseSin(1.0, 0.0, H.Grid.create(b.w, b.h, 0.0))     # amplitude 0: the body ignores the readings
    loco = helpers.locomotion(0.5)
    _, pc = oracle.run(loco.compile([H.Robot(ctrl(clean), clean)]), readings=True)
    _, pn = oracle.run(loco.compile([H.Robot(ctrl(noisy), noisy)]), readings=True)
    nt = pc["readings"].shape[1]                            # 31 ticks: t += 1/60 overshoots 0.5 by rounding
    rc, rn = pc["readings"][0].reshape(nt, 3, 4), pn["readings"][0].reshape(nt, 3, 4)   # [t, vx, vy, px]
    g1, g2 = H.JavaRandom(0), H.JavaRandom(0)
    s1 = [g1.next_gaussian() for _ in range(nt)]            # dimension-1 sensors
    s2 = [g2.next_gaussian() for _ in range(2 * nt)]        # the dimension-2 velocity sensor
    for k in range(nt):
        for v in range(3):
            assert rn[k, v, 0] - rc[k, v, 0] == pytest.approx(s1[k] * sigma * 1.0, abs=1e-12)          # Average(Touch): [0, 1]
            assert rn[k, v, 1] - rc[k, v, 1] == pytest.approx(s2[2 * k] * sigma * 1.0, abs=1e-12)      # SoftNormalization: [0, 1]
            assert rn[k, v, 2] - rc[k, v, 2] == pytest.approx(s2[2 * k + 1] * sigma * 1.0, abs=1e-12)
            assert rn[k, v, 3] - rc[k, v, 3] == pytest.approx(s1[k] * sigma * 1.0, abs=1e-12)          # Constant in [0, 1]
    assert s1[0] == pytest.approx(0.8025330637390305, abs=1e-15)   # Random(0).nextGaussian()


def test_crumpling_and_control_power():
    body = H.Grid.create(2, 1, lambda x, y: H.Voxel([H.Crumpling(), H.ControlPower(1 / 60), H.AppliedForce()]))
    r = [H.Robot(H.PhaseSin(1.0, 0.8, H.Grid.create(2, 1, 0.0)), body)]
    res, pr = oracle.run(helpers.locomotion(1.0).compile(r), readings=True, signals=True)
    rd = pr["readings"][0].reshape(-1, 2, 3)
    t = H.tick_times(1.0, 1 / 60)
    assert np.all(rd[:, :, 0] == 0.0)                       # Crumpling.java: no two masses within 0.6
    f = pr["signals"][0][:, :2]
    energy = np.cumsum(np.vstack([np.zeros((1, 2)), f[:-1] ** 2]), axis=0)   # Voxel.act adds the previous force squared
    dt = np.diff(np.concatenate([[0.0], t]))
    assert rd[:, :, 1] == pytest.approx(energy / dt[:, None], rel=1e-12, abs=1e-12)   # ControlPower.java
    assert res[0]["control_energy_last"] == pytest.approx(energy[-1].sum(), rel=1e-12)


def test_lidar_reads_the_distance_to_the_terrain():
    # Lidar.java:96-103 on flat ground (y = 5): a ray at angle a below the horizon from a centre at height h
    # reads h / -sin(a), capped at the ray length; rays pointing up read the ray length.
    body = H.Grid.create(1, 1, lambda x, y: H.Voxel([H.Lidar(10.0, -math.pi / 2, -math.pi / 4, math.pi / 2, -0.05)]))
    r = [H.Robot(H.PhaseSin(1.0, 0.0, H.Grid.create(1, 1, 0.0)), body)]
    _, pr = oracle.run(helpers.locomotion(1.0).compile(r), readings=True, trajectory=True)
    tr = pr["trajectory"][0][:, :4]
    cy = tr[:, :, 1].mean(axis=1)
    ang = (np.arctan2(tr[:, 1, 1] - tr[:, 0, 1], tr[:, 1, 0] - tr[:, 0, 0]) + np.arctan2(tr[:, 2, 1] - tr[:, 3, 1], tr[:, 2, 0] - tr[:, 3, 0])) / 2
    rd = pr["readings"][0][:, :4]
    h = cy - 5.0
    assert rd[:, 0] == pytest.approx(h / -np.sin(-math.pi / 2 + ang), rel=1e-9)
    assert rd[:, 1] == pytest.approx(h / -np.sin(-math.pi / 4 + ang), rel=1e-9)
    assert np.all(rd[:, 2] == 10.0)
    assert np.all(rd[:, 3] == 10.0)                      # 0.05 rad below the horizon: the ground is 30+ away
    # next to the left wall ((0, 100) -> (10, 5)) a ray to the west hits the wall
    body2 = H.Grid.create(1, 1, lambda x, y: H.Voxel([H.Normalization(H.Lidar(10.0, math.pi))]))
    loco = H.Locomotion(0.5, H.Locomotion.create_terrain("flat"), H.Settings(), initial_placement=10.5)
    _, p2 = oracle.run(loco.compile([H.Robot(H.PhaseSin(1.0, 0.0, H.Grid.create(1, 1, 0.0)), body2)]), readings=True, trajectory=True)
    c0 = p2["trajectory"][0][0, :4]
    cx, cyy = c0[:, 0].mean(), c0[:, 1].mean()
    wall_x = 10.0 - (cyy - 5.0) * 10.0 / 95.0            # the wall at height cy
    assert p2["readings"][0][0, 0] == pytest.approx((cx - wall_x) / 10.0, rel=1e-6)


def test_stage_clock_and_per_robot_placement():
    # a stage of a longer task (DevoLocomotion.java): the clock starts at t_start, every robot has its own placement
    robots = helpers.phase_robots(3, seed=8)
    flat = helpers.locomotion(1.0)
    a, pa = oracle.run(flat.compile(robots), trajectory=True)
    b, pb = oracle.run(flat.compile(robots, placements=[11.0, 40.0, 300.5]), trajectory=True)
    # on flat ground a robot placed elsewhere does the same thing there
    assert np.allclose(pb["trajectory"][1][:, :, 0] - 29.0, pa["trajectory"][1][:, :, 0], atol=1e-9)
    assert np.allclose(pb["trajectory"][2][:, :, 1:], pa["trajectory"][2][:, :, 1:], atol=1e-9)
    assert b["first_center_x"] - a["first_center_x"] == pytest.approx([0.0, 29.0, 289.5], abs=1e-9)
    # the clock: 1 Hz PhaseSin controllers repeat after a whole second
    c, _ = oracle.run(helpers.locomotion(3.0).compile(robots, t_start=2.0))
    assert c["n_ticks"].tolist() == [60] * 3 or c["n_ticks"].tolist() == [61] * 3
    assert c["t_first"] == pytest.approx(2.0 + 1 / 60)
    n = int(c["n_ticks"][0])
    a2, _ = oracle.run(helpers.locomotion(n / 60 - 1e-9).compile(robots))
    assert np.allclose(c["last_center_x"], a2["last_center_x"], atol=1e-6)
    # on hills the placement decides the height
    hilly = helpers.locomotion(0.2, terrain="hilly-1-10-0")
    h, ph = oracle.run(hilly.compile(robots, placements=[30.0, 100.0, 555.0]), trajectory=True)
    ys = [ph["trajectory"][i][0, :, 1].min() for i in range(3)]
    assert len({round(y, 6) for y in ys}) == 3
    with pytest.raises(ValueError):
        flat.compile(robots, placements=[1.0])


def test_dynamic_normalization_follows_the_reference_window():
    # SURVEY 8(f) rank 3, DynamicNormalization.java:26-61 over AggregatorSensor.java:56-66: readings kept while
    # t0 >= t - interval; (r - min) / (max - min) clamped to [0, 1]; 0/0 = NaN while the window holds one distinct
    # value -- always at the first tick (the reference's own behaviour, reproduced, not repaired).
    mk_raw = lambda x, y: H.Voxel([H.AreaRatio(), H.Velocity(True, 8.0, H.Velocity.X, H.Velocity.Y)])
    mk_dyn = lambda x, y: H.Voxel([H.DynamicNormalization(H.AreaRatio(), 0.25),
                                   H.SoftNormalization(H.DynamicNormalization(H.Velocity(True, 8.0, H.Velocity.X, H.Velocity.Y), 0.5))])
    ctrl = lambda: H.PhaseSin(1.0, 1.0, H.Grid.create(3, 2, lambda x, y: 0.7 * x + 0.2 * y))
    loco = helpers.locomotion(2.0)
    _, raw = oracle.run(loco.compile([H.Robot(ctrl(), H.Grid.create(3, 2, mk_raw))]), readings=True)
    _, dyn = oracle.run(loco.compile([H.Robot(ctrl(), H.Grid.create(3, 2, mk_dyn))]), readings=True)
    t = H.tick_times(2.0, 1 / 60)
    nt = len(t)
    raw, dyn = raw["readings"][0].reshape(nt, 6, 3), dyn["readings"][0].reshape(nt, 6, 3)
    assert [d.min for d in mk_dyn(0, 0).get_sensors()[0].domains()] == [0.0] and \
        [d.max for d in mk_dyn(0, 0).get_sensors()[1].domains()] == [1.0, 1.0]
    for ch, interval, soft in ((0, 0.25, False), (1, 0.5, True), (2, 0.5, True)):
        for k in range(nt):
            keep = [q for q in range(k + 1) if not t[q] < t[k] - interval]
            keep = list(range(keep[0], k + 1))                 # the TreeMap drops from the oldest key on
            mn, mx = raw[keep, :, ch].min(0), raw[keep, :, ch].max(0)
            with np.errstate(invalid="ignore", divide="ignore"):
                want = np.clip((raw[k, :, ch] - mn) / (mx - mn), 0.0, 1.0)
            if soft:
                want = np.tanh((want * 2 - 1) * 2) / 2 + 0.5
            if k == 0:
                assert np.all(np.isnan(dyn[0, :, ch]))         # one reading in the window: 0/0
            else:
                assert np.array_equal(np.isnan(want), np.isnan(dyn[k, :, ch]))
                ok = ~np.isnan(want)
                assert dyn[k, ok, ch] == pytest.approx(want[ok], abs=1e-12)
    assert np.isfinite(dyn[1:]).mean() > 0.99
    # a controller that reads the NaN of the first tick asks for applyForce(NaN): neither branch of Voxel.java:230-235
    # runs, the springs keep their rest distance, lastAppliedForce (and with it the control energy) is NaN; from the
    # second tick on the window holds two readings and the loop is closed with finite numbers
    body = H.Grid.create(2, 1, lambda x, y: H.Voxel([H.DynamicNormalization(H.AreaRatio(), 0.25)]))
    nin, nout = H.CentralizedSensing.n_of_inputs(body), H.CentralizedSensing.n_of_outputs(body)
    mlp = H.MultiLayerPerceptron("TANH", nin, [], nout, np.full(H.MultiLayerPerceptron.count_weights([nin, nout]), 0.1))
    res, pr = oracle.run(helpers.locomotion(0.5).compile([H.Robot(H.CentralizedSensing(body, mlp), body)]), signals=True, trajectory=True)
    assert res[0]["status"] == 0 and np.isfinite(pr["trajectory"]).all()
    assert np.isnan(pr["signals"][0, 0, :2]).all() and np.isfinite(pr["signals"][0, 5:, :2]).all()
    assert np.isnan(res[0]["control_energy_last"])


def _rope_material(lo=0.9, hi=1.1):
    m = H.abi.default_material()
    m.area_ratio_passive_min, m.area_ratio_passive_max = lo, hi
    return m


def _spring_lengths(tr, half=0.525):
    """[tick][voxel][4]: lengths of the side-parallel internal springs 0..3 (Voxel.java:305-330) from a trajectory."""
    h = half
    a1 = np.array([[+h, -h], [-h, -h], [-h, +h], [+h, +h]])
    a2 = np.array([[-h, -h], [-h, +h], [+h, +h], [+h, -h]])
    nt, nb = tr.shape[:2]
    b = tr.reshape(nt, nb // 4, 4, 6)
    out = np.zeros((nt, nb // 4, 4))
    for s in range(4):
        p, q = b[:, :, s], b[:, :, (s + 1) % 4]
        def world(m, a):
            c, sn = np.cos(m[..., 2]), np.sin(m[..., 2])
            return m[..., 0] + c * a[0] - sn * a[1], m[..., 1] + sn * a[0] + c * a[1]
        x1, y1 = world(p, a1[s])
        x2, y2 = world(q, a2[s])
        out[:, :, s] = np.hypot(x1 - x2, y1 - y2)
    return out


def test_rope_limits_hold_the_passive_range():
    # north_star "rope forces", SURVEY 8(f) rank 4: a finite areaRatioPassiveRange turns into lower / upper limits of
    # every DistanceJoint (Voxel.java:262-285,331-338).  With the passive range (0.9, 1.1) inside the active one
    # (0.8, 1.2) a full-amplitude controller drives the springs into their limits: they hold (one-sided rigid rows +
    # position correction), where the unlimited voxel goes well beyond.
    L, ms = 3.0, 1.05
    lo, hi = np.sqrt(L * L * 0.9) - 2 * ms, np.sqrt(L * L * 1.1) - 2 * ms
    ctrl = lambda: H.PhaseSin(1.0, 1.0, H.Grid.create(2, 1, 0.0))
    runs = {}
    for name, mat in (("free", None), ("rope", _rope_material())):
        body = H.Grid.create(2, 1, lambda x, y: H.Voxel([], mat))
        _, pr = oracle.run(helpers.locomotion(3.0).compile([H.Robot(ctrl(), body)]), trajectory=True)
        runs[name] = _spring_lengths(pr["trajectory"][0])
    assert runs["free"][0] == pytest.approx(L - 2 * ms, abs=1e-12)
    assert runs["free"].min() < lo - 0.03 and runs["free"].max() > hi + 0.03       # the actuation asks for more
    assert runs["rope"].min() > lo - 0.011 and runs["rope"].max() < hi + 0.011     # ... the limits hold (2 x linearTolerance)
    assert runs["rope"].min() < lo + 0.01 and runs["rope"].max() > hi - 0.01       # ... and are reached
    # one-sided: only the upper limit
    body = H.Grid.create(2, 1, lambda x, y: H.Voxel([], _rope_material(-np.inf, 1.1)))
    _, pr = oracle.run(helpers.locomotion(3.0).compile([H.Robot(ctrl(), body)]), trajectory=True)
    up = _spring_lengths(pr["trajectory"][0])
    assert up.max() < hi + 0.011 and up.min() < lo - 0.03
    # SpringRange validation (Voxel.java:189-193): a passive range that excludes the rest length
    with pytest.raises(ValueError, match="Wrong spring range"):
        H.Voxel([], _rope_material(1.05, 1.3))


def _breakable_restatement(bv, t, forces_in, area_ratio, raw_readings, rd_domains, read_by_controller):
    """BreakableVoxel.java:127-193 restated in Python for ONE voxel over the ticks `t`: given what the controller asks
    for (forces_in[k], or a callable of the effective readings), the voxel's area ratio and raw readings per tick ->
    (applied force, isBroken at sense time, effective readings) per tick."""
    rnd = H.JavaRandom(bv.random_seed)
    comps, trigs, types = bv.COMPONENTS, bv.TRIGGERS, bv.TYPES
    state = {c: "NONE" for c in comps}
    cnt = {g: 0.0 for g in trigs}
    last_t = last_break = last_ce = last_are = 0.0
    ce = are = 0.0
    last_f = 0.0
    snap = None
    out_f, out_broken, out_eff = [], [], []
    for k in range(len(t)):
        # Voxel.act: energies, then the sensors (Malfunction reads the state BEFORE this tick's update)
        are += area_ratio[k] ** 2
        ce += last_f ** 2
        out_broken.append(any(s != "NONE" for s in state.values()))
        if state["SENSORS"] == "NONE" or snap is None:
            snap = np.array(raw_readings[k], dtype=float)
        cnt["TIME"] += t[k] - last_t
        cnt["CONTROL"] += ce - last_ce
        cnt["AREA"] += are - last_are
        last_t, last_ce, last_are = t[k], ce, are
        breaking = False
        for g in trigs:
            if g not in bv.trigger_thresholds:
                continue
            thr = bv.trigger_thresholds[g]
            with np.errstate(divide="ignore", invalid="ignore"):
                p = 1.0 - np.tanh(np.float64(thr) / np.float64(cnt[g]))
            if rnd.next_double() < p:
                cnt = {q: 0.0 for q in trigs}
                keys = [c for c in comps if c in bv.malfunctions]
                if keys:
                    breaking = True
                    c = keys[rnd.next_int(len(keys))]
                    ts = [q for q in types if q in bv.malfunctions[c]]
                    state[c] = ts[rnd.next_int(len(ts))]
        if breaking:
            last_break = t[k]
        if t[k] - last_break > bv.restore_time:
            state = {c: "NONE" for c in comps}
        # controller.control: getSensorReadings (if it reads), then applyForce
        eff = snap
        if read_by_controller:
            if state["SENSORS"] == "ZERO":
                eff = np.zeros_like(snap)
            elif state["SENSORS"] == "RANDOM":
                eff = np.array([rnd.next_double() * (hi - lo) + lo for lo, hi in rd_domains])
        out_eff.append(np.array(eff))
        f = forces_in(k, eff) if callable(forces_in) else forces_in[k]
        if state["ACTUATOR"] == "ZERO":
            f = 0.0
        elif state["ACTUATOR"] == "FROZEN":
            f = last_f
        elif state["ACTUATOR"] == "RANDOM":
            f = rnd.next_double() * 2.0 - 1.0
        if abs(f) > 1:
            f = math.copysign(1.0, f)
        last_f = f
        out_f.append(f)
    return np.array(out_f), np.array(out_broken), out_eff


def test_breakable_voxel_actuator_and_triggers_follow_the_reference():
    # SURVEY 8(f) rank 4, BreakableVoxel.java: all three triggers, all three actuator malfunctions, restore -- the
    # oracle's applied forces and Malfunction readings against a Python restatement driven by the same java.util.Random
    t = H.tick_times(6.0, 1 / 60)
    mk = lambda x, y: H.BreakableVoxel([H.AreaRatio(), H.Malfunction()], 7 + 13 * x, {"ACTUATOR": {"ZERO", "FROZEN", "RANDOM"}},
                                       {"TIME": 1.5, "CONTROL": 25.0, "AREA": 60.0}, 0.4)
    body = H.Grid.create(3, 1, mk)
    ctrl = H.PhaseSin(1.0, 0.9, H.Grid.create(3, 1, lambda x, y: 0.8 * x))
    res, pr = oracle.run(helpers.locomotion(6.0).compile([H.Robot(ctrl, body)]), signals=True, readings=True)
    assert res[0]["status"] == 0
    rd = pr["readings"][0].reshape(len(t), 3, 2)
    seen = set()
    for j in range(3):
        want = np.sin(2 * np.pi * 1.0 * t + 0.8 * j) * 0.9
        f, broken, _ = _breakable_restatement(body.get(j, 0), t, want, rd[:, j, 0], rd[:, j, :], [(0, 1)] * 2, False)
        assert pr["signals"][0, :, j] == pytest.approx(f, abs=1e-12)
        assert np.array_equal(rd[:, j, 1] == 1.0, broken)
        seen |= {"broken"} if broken.any() else set()
        assert broken.any() and not broken.all()                      # it breaks and it is restored
        assert np.abs(f - want).max() > 0.1                           # ... and the malfunction changes what is applied
    # a plain voxel in the same grid is left alone, and "broken-1-0" freezes every actuator from the first tick on
    robot = H.Robot(ctrl, H.Grid.create(3, 1, lambda x, y: H.Voxel([H.AreaRatio(), H.Malfunction()])))
    broken_robot = H.RobotUtils.build_robot_transformation("broken-1-0")(robot)
    assert all(isinstance(v, H.BreakableVoxel) for v in broken_robot.get_voxels().values())
    _, pb = oracle.run(helpers.locomotion(2.0).compile([broken_robot]), signals=True, readings=True)
    assert np.all(pb["signals"][0, :, :3] == 0.0)                     # FROZEN at tick 0: the last applied force is 0
    assert np.all(pb["readings"][0].reshape(-1, 3, 2)[1:, :, 1] == 1.0)
    half = H.RobotUtils.build_robot_transformation("broken-0.5-3")(robot)
    kinds = [isinstance(v, H.BreakableVoxel) for v in half.get_voxels().values()]
    rnd = H.JavaRandom(3)
    want_kinds = []
    for _ in range(3):
        b = not (rnd.next_double() > 0.5)
        want_kinds.append(b)
        if b:
            rnd.next_int()
    assert kinds == want_kinds
    br = H.RobotUtils.build_robot_transformation("breakable-time-2/0.5-1/0.1-5")(robot)
    v0 = br.get_voxels().get(0, 0)
    rnd = H.JavaRandom(5)
    assert v0.random_seed == rnd.next_int() and v0.trigger_thresholds == {"TIME": rnd.next_gaussian() * 0.5 + 2.0}
    assert v0.restore_time == rnd.next_gaussian() * 0.1 + 1.0 and v0.malfunctions == {"ACTUATOR": frozenset({"FROZEN"})}
    with pytest.raises(ValueError, match="Unknown transformation"):
        H.RobotUtils.build_robot_transformation("melted-1")
    with pytest.raises(ValueError, match="Unsupported structure"):
        H.BreakableVoxel([], 0, {"STRUCTURE": {"RANDOM"}}, {"TIME": 1.0}, 1.0)


def test_breakable_voxel_sensors_and_structure():
    # SENSORS malfunctions reach the controller (CentralizedSensing reads Voxel.getSensorReadings): ZERO, FROZEN, RANDOM
    # (uniform in each reading's domain, drawn from the voxel's own generator at control time)
    t = H.tick_times(4.0, 1 / 60)
    mk = lambda x, y: H.BreakableVoxel([H.SoftNormalization(H.AreaRatio()), H.Velocity(True, 8.0, H.Velocity.X)], 100 + x,
                                       {"SENSORS": {"ZERO", "FROZEN", "RANDOM"}}, {"TIME": 0.8}, 0.3)
    body = H.Grid.create(2, 1, mk)
    nin, nout = H.CentralizedSensing.n_of_inputs(body), H.CentralizedSensing.n_of_outputs(body)
    w = np.random.default_rng(3).uniform(-1, 1, size=H.MultiLayerPerceptron.count_weights([nin, nout]))
    mlp = H.MultiLayerPerceptron("TANH", nin, [], nout, w)
    res, pr = oracle.run(helpers.locomotion(4.0).compile([H.Robot(H.CentralizedSensing(body, mlp), body)]), signals=True, readings=True)
    rd = pr["readings"][0].reshape(len(t), 2, 2)
    doms = [(0.0, 1.0), (-8.0, 8.0)]
    effs = []
    for j in range(2):
        _, _, eff = _breakable_restatement(body.get(j, 0), t, np.zeros(len(t)), np.ones(len(t)), rd[:, j, :], doms, True)
        effs.append(np.stack(eff))
    changed = 0
    for k in range(len(t)):
        x = np.concatenate([effs[0][k], effs[1][k]])
        want = np.clip(mlp.apply(x), -1, 1)
        assert pr["signals"][0, k, :2] == pytest.approx(want, abs=1e-12), k
        changed += int(np.abs(x - rd[k].reshape(-1)).max() > 1e-9)
    assert 10 < changed < len(t) - 10
    # STRUCTURE FROZEN: the voxel's springs turn into rigid distance joints at the break (setFrequency(0): they now hold
    # the commanded rest distance exactly) and -- as in the reference, which re-applies the structure state at a break
    # only -- stay rigid after the restore
    mk2 = lambda x, y: H.BreakableVoxel([], 1, {"STRUCTURE": {"FROZEN"}}, {"TIME": 0.5}, 0.2)
    body2 = H.Grid.create(1, 1, mk2)
    rob = H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(1, 1, 0.0)), body2)
    _, p2 = oracle.run(helpers.locomotion(4.0).compile([rob]), trajectory=True)
    ln = _spring_lengths(p2["trajectory"][0])[:, 0, 0]
    t4 = H.tick_times(4.0, 1 / 60)
    _, broken, _ = _breakable_restatement(body2.get(0, 0), t4, np.zeros(len(t4)), np.ones(len(t4)), np.zeros((len(t4), 0)), [], False)
    k0 = int(np.argmax(broken)) - 1                                 # the tick of the first break
    assert 10 < k0 < 150 and not broken[k0 + 1:].all()              # ... and there are restored stretches afterwards
    # the rest distance a rigid joint is held at: what applyForce set at the previous control step (Voxel.java:230-235)
    f = np.concatenate([[0.0], np.sin(2 * np.pi * t4[:-1])])
    lo, rest, hi = np.sqrt(9 * 0.8) - 2.1, 0.9, np.sqrt(9 * 1.2) - 2.1
    cmd = np.where(f >= 0, rest - (rest - lo) * f, rest + (hi - rest) * -f)
    assert np.abs(ln - cmd)[1:k0].max() > 0.02                      # soft at first: the length lags the command
    assert np.abs(ln - cmd)[k0 + 3:].max() < 0.006                  # rigid from the break on (within linearTolerance), restore or not


def _soft_heavy_material():
    m = H.abi.default_material()
    m.spring_f, m.mass, m.friction, m.side_length = 5.0, 2.0, 3.0, 2.5
    return m


def test_materials_per_shape_class_in_one_batch():
    # Voxel.java:104-118: the 13 constructor fields belong to the voxel.  The device takes one material per ROBOT;
    # robots of different materials share a batch (vsr_shape.material): each gives what it gives alone.
    mats = [None, _soft_heavy_material()]
    robots = []
    for i in range(4):
        body = H.Grid.create(3, 2, lambda x, y, m=mats[i % 2]: H.Voxel([H.AreaRatio()], m))
        robots.append(H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(3, 2, lambda x, y: 0.5 * x + 0.1 * i)), body))
    loco = helpers.locomotion(3.0)
    bd = loco.compile(robots)
    assert bd.desc.n_shapes == 4 or bd.desc.n_shapes == 2
    together, _ = oracle.run(bd)
    for i, r in enumerate(robots):
        alone, _ = oracle.run(loco.compile([r]))
        assert alone[0].tobytes() == together[i].tobytes(), i
    assert abs(together[0]["last_center_x"] - together[1]["last_center_x"]) > 1e-3   # the material matters
    # ... but not two materials inside one robot
    mixed = H.Grid.create(2, 1, lambda x, y: H.Voxel([], mats[x]))
    with pytest.raises(NotImplementedError, match="one voxel material per robot"):
        loco.compile([H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(2, 1, 0.0)), mixed)])


def test_at_rest_sleeping_never_touches_an_actuated_robot():
    # SURVEY C.7, the one dyn4j feature the kernel leaves out: at-rest detection.  An island sleeps after 0.5 s below
    # 0.01 u/s and 2 deg/s and wakes when DistanceJoint.setRestDistance changes a value.  A controller changes the rest
    # distances at every tick, so the switch changes NOTHING for actuated robots (bitwise, configs 2 and 3) ...
    import workloads
    for loco, robots in (workloads.config2(24, seed=0), workloads.config3(8, hidden=(21, 21), seed=1)):
        bd = H.Locomotion(6.0, loco.ground_profile, loco.settings).compile(robots)
        a, _ = oracle.run(bd, n_threads=8)
        b2, _ = oracle.run(bd, n_threads=8, sleeping=True)
        assert a.tobytes() == b2.tobytes()
    # ... and only a passive robot ever sleeps: it settles, its velocities are zeroed for good, and it stays where the
    # always-awake robot keeps creeping by less than a tenth of the linear tolerance
    bd = helpers.locomotion(8.0).compile(helpers.phase_robots(1, shape="box-2x1", amp=0.0))
    _, awake = oracle.run(bd, trajectory=True)
    _, slept = oracle.run(bd, trajectory=True, sleeping=True)
    v = np.abs(slept["trajectory"][0][..., 3:]).max(axis=1).max(axis=1)
    k0 = int(np.argmax(v == 0.0))
    assert 60 < k0 < 420 and np.all(v[k0:] == 0.0)                    # asleep from k0 on
    assert np.abs(slept["trajectory"][0][k0:, :, :2] - slept["trajectory"][0][k0, :, :2]).max() == 0.0
    assert np.abs(awake["trajectory"][0][-1, :, :2] - slept["trajectory"][0][-1, :, :2]).max() < 5e-4
