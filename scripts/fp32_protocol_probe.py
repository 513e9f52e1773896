"""Exploration for the SURVEY 8(c) fp32 parity protocol: kernel fp32 vs oracle (canonical + K shuffles)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
import workloads, helpers
from oracle import oracle
from scipy.stats import spearmanr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = os.cpu_count()

def disp(r): return r["last_center_x"] - r["first_center_x"]

def crit(d, ref, sig):
    tol = np.maximum(np.maximum(3 * sig, 0.05 * np.abs(ref)), 0.5)
    return float((np.abs(d - ref) <= tol).mean())

for name, (loco, robots) in (("cfg2", workloads.config2(N)), ("cfg3[21,21]", workloads.config3(N, (21, 21))),
                             ("cfg3[]", workloads.config3(N, ())), ("cfg4", workloads.config4(N))):
    bd = loco.compile(robots)
    t0 = time.time()
    with H.DeviceBatch(bd) as db:
        g = db.upload(0).run().results()
    bd64 = loco.compile(robots, precision=H.abi.VSR_PRECISION_F64)
    with H.DeviceBatch(bd64) as db:
        g64 = db.upload(0).run().results()
    t1 = time.time()
    o, _ = oracle.run(bd, n_threads=T)
    t2 = time.time()
    sh = [disp(oracle.run(bd, order=2, seed=100 + k, n_threads=T)[0]) for k in range(K + 1)]
    t3 = time.time()
    dg, dg64, do = disp(g), disp(g64), disp(o)
    sig = np.std(np.stack(sh[:K]), axis=0, ddof=1)
    print(f"== {name}: N={N} gpu {t1-t0:.1f}s oracle {t2-t1:.1f}s shuffles {t3-t2:.1f}s status!=0: f32 {(g['status']!=0).sum()} f64 {(g64['status']!=0).sum()} oracle {(o['status']!=0).sum()}")
    print(f"   sigma_perm: median {np.median(sig):.3f} mean {sig.mean():.3f} p90 {np.percentile(sig, 90):.3f}")
    print(f"   within tol: f32 {crit(dg, do, sig):.3f}  f64-kernel {crit(dg64, do, sig):.3f}  extra shuffle {crit(sh[K], do, sig):.3f}")
    print(f"   mean d: f32 {dg.mean():.3f} f64k {dg64.mean():.3f} oracle {do.mean():.3f} shuffle {sh[K].mean():.3f} (rel f32 {(dg.mean()/do.mean()-1)*100:.2f}% shuffle {(sh[K].mean()/do.mean()-1)*100:.2f}%)")
    print(f"   median d: f32 {np.median(dg):.3f} oracle {np.median(do):.3f} shuffle {np.median(sh[K]):.3f} (rel f32 {(np.median(dg)/np.median(do)-1)*100:.2f}% shuffle {(np.median(sh[K])/np.median(do)-1)*100:.2f}%)")
    print(f"   spearman: f32 {spearmanr(dg, do)[0]:.3f} f64k {spearmanr(dg64, do)[0]:.3f} shuffle {spearmanr(sh[K], do)[0]:.3f}; |d f64k - oracle| max {np.abs(dg64-do).max():.3e}")

# short horizon: per-tick max |dpos| curves
for name, (loco, robots) in (("cfg2", workloads.config2(256)), ("cfg3[21,21]", workloads.config3(256, (21, 21), final_t=2.0)),
                             ("cfg4", workloads.config4(128, final_t=2.0))):
    loco2 = H.Locomotion(2.0, loco.ground_profile, loco.settings)
    bd = loco2.compile(robots, precision=H.abi.VSR_PRECISION_F32, trajectory_robots=len(robots))
    with H.DeviceBatch(bd) as db:
        db.upload(0).run()
        tr = [db.trajectory(i) for i in range(len(robots))]
    _, pr = oracle.run(bd, trajectory=True, n_threads=T)
    err = np.zeros((len(robots), bd.n_ticks))
    for i in range(len(robots)):
        nb = tr[i].shape[1]
        err[i] = np.abs(tr[i][:, :, :2] - pr["trajectory"][i][:, :nb, :2]).max(axis=(1, 2))
    cum = np.maximum.accumulate(err, axis=1)
    for k in (4, 9, 24, 39, 59, 89, 119):
        c = cum[:, k]
        print(f"   {name} ticks 1..{k+1}: median {np.median(c):.2e} p90 {np.percentile(c, 90):.2e} p99 {np.percentile(c, 99):.2e} max {c.max():.2e}  frac>1e-4 {(c>1e-4).mean():.3f} frac>1e-2 {(c>1e-2).mean():.3f}")
