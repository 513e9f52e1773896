"""The fp32 PRODUCT path against the fp64 oracle: the parity protocol of SURVEY.md 8(c), on BASELINE configs 2, 3, 4.

The dynamics is chaotic (stick-slip contacts): the oracle's OWN 20 s displacement of a robot moves by sigma_perm
(median 5.4 units on config 2) when nothing but the Gauss-Seidel constraint order changes -- a change north_star
concedes.  SURVEY 8(c) therefore asks for its numbers "to be calibrated from the oracle's own sensitivity".  The
calibration is in tests/golden/fp32_protocol.npz (made by tests/golden/make_fp32_protocol.py, re-checked against the
live oracle by tests/test_host_golden.py): for N = 1024 robots per config the canonical-order oracle and K + 1 = 9
seeded shuffled orders.  Every statistic of "fp32 kernel vs canonical oracle" is compared with the SAME statistic of
"shuffled fp64 oracle vs canonical oracle" (the null distribution, K + 1 leave-one-out samples):

 (i)   per robot: |d_fp32 - d_oracle| <= max(3 sigma_perm, 5 % |d_oracle|, 0.5) for at least as many robots as the
       worst reordered oracle run reaches, minus 0.02 (the survey's flat 95 % is not reached by the oracle against
       itself: with sigma estimated from 8 samples the criterion is a t-test at ~93 %);
 (ii)  population mean and median displacement (= velocity x 19.98 s) within max(2 %, 1.5 x the largest deviation a
       reordered oracle shows) of the canonical oracle's;
 (iii) Spearman rank correlation of the per-robot displacements no lower than the worst reordered run's minus 0.03.

Short horizon (ticks 1..120): the per-tick max |dpos| curve is printed and bounded, next to the curve of the fp64
oracle whose STATE is rounded to float after every step (what fp32 storage alone costs): the kernel must stay within
3x of that curve (measured: at or below it).
"""
import importlib
import os
import sys

import numpy as np
import pytest
from scipy.stats import spearmanr

H = importlib.import_module("2dhmsr_b200")
A = H.abi
from oracle import oracle
import helpers

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_fp32_protocol as proto

pytestmark = pytest.mark.gpu

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "fp32_protocol.npz"))


def _disp(r):
    return r["last_center_x"] - r["first_center_x"]


def _within(d, canon, sig):
    tol = np.maximum(np.maximum(3.0 * sig, 0.05 * np.abs(canon)), 0.5)   # max(3 sigma_perm, 5 % |d|, 0.5)
    return float((np.abs(d - canon) <= tol).mean())


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_fp32_full_episode_protocol(cfg):
    loco, robots = proto.configs()[cfg]
    assert len(robots) == proto.N >= 1024
    bd = loco.compile(robots)                                   # fp32, the product path
    assert bd.n_ticks == 1200
    res, _ = helpers.gpu_run(bd)
    assert np.all(res["status"] == 0), np.bincount(res["status"])
    d = _disp(res)
    canon, shuf = FIX[cfg + "_canon"], FIX[cfg + "_shuf"]
    K = shuf.shape[0] - 1
    sig = np.std(shuf[:K], axis=0, ddof=1)                      # sigma_perm per robot, K = 8 shuffled orders
    frac = _within(d, canon, sig)
    null_frac, null_mean, null_med, null_rho = [], [], [], []
    for k in range(K + 1):                                      # leave-one-out: shuffle k judged by the other K
        sig_k = np.std(np.delete(shuf, k, axis=0), axis=0, ddof=1)
        null_frac.append(_within(shuf[k], canon, sig_k))
        null_mean.append(abs(shuf[k].mean() - canon.mean()))
        null_med.append(abs(np.median(shuf[k]) - np.median(canon)))
        null_rho.append(spearmanr(shuf[k], canon)[0])
    rho = spearmanr(d, canon)[0]
    dm, dmed = abs(d.mean() - canon.mean()), abs(np.median(d) - np.median(canon))
    tol_mean = max(0.02 * abs(canon.mean()), 1.5 * max(null_mean))
    tol_med = max(0.02 * abs(np.median(canon)), 1.5 * max(null_med))
    print(f"\n[{cfg}] N={len(d)} sigma_perm median {np.median(sig):.3f}; within tol: fp32 {frac:.3f} "
          f"(reordered oracle {min(null_frac):.3f}..{max(null_frac):.3f}; survey target 0.95); "
          f"mean d fp32 {d.mean():.3f} oracle {canon.mean():.3f} (|diff| {dm:.3f} = {100 * dm / abs(canon.mean()):.1f} %, "
          f"reordered up to {max(null_mean):.3f}); median fp32 {np.median(d):.3f} oracle {np.median(canon):.3f} "
          f"(|diff| {dmed:.3f}, reordered up to {max(null_med):.3f}); Spearman fp32 {rho:.3f} "
          f"(reordered {min(null_rho):.3f}..{max(null_rho):.3f})")
    assert frac >= min(0.95, min(null_frac) - 0.02)             # (i)
    assert dm <= tol_mean                                       # (ii) mean: max(2 %, 1.5 x reordered oracle)
    assert dmed <= tol_med                                      # (ii) median
    assert rho >= min(null_rho) - 0.03                          # (iii)
    # control energy is a pure function of the inputs for the open-loop configs: tight
    if cfg != "cfg3":
        ores, _ = oracle.run(bd, n_threads=os.cpu_count(), robot_end=32)
        assert np.allclose(res["control_energy_last"][:32], ores["control_energy_last"], rtol=1e-5)


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3"])
def test_fp32_full_episode_protocol_without_continuous_detection(cfg):
    """The same protocol under ContinuousDetectionMode.NONE (vsr_settings.continuous_detection = 0, the MODE 1 kernel):
    fp32 kernel against the fixture's canonical oracle run without the pass.  The null distribution (what a reshuffle
    of the solver order does) is the fixture's: the pass does not change how chaotic a robot is."""
    loco, robots = proto.configs()[cfg]
    st = H.Settings()
    for f, _t in st.raw()._fields_:
        setattr(st.raw(), f, getattr(loco.settings.raw(), f))
    bd = H.Locomotion(20.0, loco.ground_profile, st.set(continuous_detection=0)).compile(robots)
    res, _ = helpers.gpu_run(bd)
    assert np.all(res["status"] == 0)
    d, canon, on, shuf = _disp(res), FIX[cfg + "_canon_nopass"], FIX[cfg + "_canon"], FIX[cfg + "_shuf"]
    K = shuf.shape[0] - 1
    sig = np.std(shuf[:K], axis=0, ddof=1)
    null_frac, null_mean, null_rho = [], [], []
    for k in range(K + 1):
        sig_k = np.std(np.delete(shuf, k, axis=0), axis=0, ddof=1)
        null_frac.append(_within(shuf[k], on, sig_k))
        null_mean.append(abs(shuf[k].mean() - on.mean()))
        null_rho.append(spearmanr(shuf[k], on)[0])
    frac, rho, dm = _within(d, canon, sig), spearmanr(d, canon)[0], abs(d.mean() - canon.mean())
    print(f"\n[{cfg}, no continuous detection] mean d fp32 {d.mean():.3f} oracle {canon.mean():.3f} (with the pass "
          f"{on.mean():.3f}); within tol {frac:.3f} (reordered oracle {min(null_frac):.3f}); Spearman {rho:.3f} "
          f"(reordered {min(null_rho):.3f})")
    assert d.mean() > on.mean() + 0.5 * max(null_mean)           # the switch is live in the kernel
    assert frac >= min(0.95, min(null_frac) - 0.03)
    assert dm <= max(0.03 * abs(canon.mean()), 1.5 * max(null_mean))
    assert rho >= min(null_rho) - 0.04


def _curves(loco, robots, n_ticks=120):
    """max |dpos| over bodies, cumulative over ticks, per robot: fp32 kernel vs oracle, and phase-perturbed oracle vs oracle."""
    loco2 = H.Locomotion(2.0, loco.ground_profile, loco.settings)
    bd = loco2.compile(robots, precision=A.VSR_PRECISION_F32, trajectory_robots=len(robots))
    with H.DeviceBatch(bd) as db:
        db.upload(0).run()
        tr = [db.trajectory(i) for i in range(len(robots))]
    _, pr = oracle.run(bd, trajectory=True, n_threads=os.cpu_count())
    err = np.zeros((len(robots), n_ticks))
    for i in range(len(robots)):
        nb = tr[i].shape[1]
        err[i] = np.abs(tr[i][:n_ticks, :, :2] - pr["trajectory"][i][:n_ticks, :nb, :2]).max(axis=(1, 2))
    return np.maximum.accumulate(err, axis=1), pr["trajectory"]


def test_fp32_short_horizon_curve_config2():
    import workloads
    n = 256
    loco, robots = workloads.config2(n, seed=0)
    cum, otr = _curves(loco, robots)
    # the yardstick: the fp64 oracle whose body state is ROUNDED TO FLOAT after every step (vsr_oracle_opts.state_fp32)
    # -- what fp32 storage of the state alone does to a trajectory, all arithmetic still in double
    bdq = H.Locomotion(2.0, loco.ground_profile, loco.settings).compile(robots)
    _, pp = oracle.run(bdq, trajectory=True, n_threads=os.cpu_count(), state_fp32=True)
    perr = np.maximum.accumulate(np.abs(pp["trajectory"][:, :120, :, :2] - otr[:, :120, :, :2]).max(axis=(2, 3)), axis=1)
    print()
    for k in (4, 9, 24, 39, 59, 89, 119):
        c, p = cum[:, k], perr[:, k]
        print(f"   ticks 1..{k + 1:3d}: fp32 median {np.median(c):.2e} p90 {np.percentile(c, 90):.2e} max {c.max():.2e} "
              f"frac>1e-2 {(c > 1e-2).mean():.3f} | oracle(fp32 state) median {np.median(p):.2e} p90 {np.percentile(p, 90):.2e} "
              f"max {p.max():.2e}")
    # ticks 1..5: rounding only (free fall; corner neighbours start to overlap: a contact that exists in one
    # precision and not yet in the other -- narrowphase margin 1e-12 vs 1e-6 -- is the only discrete event)
    assert np.median(cum[:, 4]) < 1e-5 and cum[:, 4].max() < 2e-3
    # ticks 1..25 (SURVEY proposed 1e-4: holds for the median robot up to tick ~15; a position ulp of 1e-6 under springs of
    # k = 316 is a force noise of 3e-4 per spring and tick, which integrates to ~2e-4 by tick 25 -- the fp64 oracle with
    # an fp32-rounded state shows 5.5e-4): median < 1e-3, and
    # at most 2 % of the robots past 1e-2 (a flipped discrete decision: contact on/off, position-iteration count)
    assert np.median(cum[:, 24]) < 1e-3 and (cum[:, 24] > 1e-2).mean() <= 0.02 and cum[:, 24].max() < 0.05
    # ticks 1..120 (first impact and two bounces): SURVEY's 1e-2 holds for the median robot
    assert np.median(cum[:, 119]) < 1e-2 and np.percentile(cum[:, 119], 90) < 0.3 and cum[:, 119].max() < 2.0
    # ... and none of this is the kernel's doing: it stays within 3x of what rounding the STATE to fp32 does to the
    # fp64 oracle itself (measured: the kernel is at or below that curve)
    for k in (24, 59, 119):
        assert np.median(cum[:, k]) <= 3 * np.median(perr[:, k]), k
        assert np.percentile(cum[:, k], 90) <= 3 * np.percentile(perr[:, k], 90), k


@pytest.mark.parametrize("cfg", ["cfg3", "cfg4"])
def test_fp32_short_horizon_curve_closed_loop_and_hills(cfg):
    import workloads
    loco, robots = (workloads.config3(256, (21, 21), seed=1) if cfg == "cfg3" else workloads.config4(128, seed=2))
    cum, _ = _curves(loco, robots)
    print()
    for k in (4, 24, 59, 119):
        c = cum[:, k]
        print(f"   {cfg} ticks 1..{k + 1:3d}: fp32 median {np.median(c):.2e} p90 {np.percentile(c, 90):.2e} max {c.max():.2e}")
    if cfg == "cfg3":
        # closed loop: a Touch reading that flips one tick earlier moves an MLP input by 1/15 -> the controller output
        # changes at once; stated bounds: median 5e-2 by tick 25, 1.5 units by tick 120
        assert np.median(cum[:, 24]) < 5e-2 and np.median(cum[:, 119]) < 1.5 and cum[:, 119].max() < 6.0
    else:
        assert np.median(cum[:, 24]) < 5e-3 and np.median(cum[:, 119]) < 0.3 and cum[:, 119].max() < 2.0
