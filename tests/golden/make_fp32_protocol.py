"""Oracle half of the fp32 parity protocol, precomputed: tests/golden/fp32_protocol.npz.

For each of BASELINE configs 2, 3 ([21, 21] closed loop) and 4, N = 1024 seeded robots, full 20 s episodes:
  canon[i]      20 s displacement of robot i, oracle (fp64) in the canonical Gauss-Seidel order (= the kernel's),
                default settings (continuous detection on, as dyn4j's `new Settings()`)
  canon_nopass  the same with vsr_settings.continuous_detection = 0
  shuf[k][i]    the same under K + 1 = 9 seeded random constraint orders (VSR_ORACLE_ORDER_SHUFFLED)
The shuffles measure the INTRINSIC spread of a robot's outcome under a change north_star concedes (solver order):
sigma_perm[i] = std_k shuf[k][i].  tests/test_fp32_protocol_gpu.py compares the fp32 kernel against these;
tests/test_host_golden.py re-runs a slice of the oracle against the file so that it cannot go stale.

    python tests/golden/make_fp32_protocol.py        (about 20 minutes on 8 cores)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

N, K = 1024, 8
SHUFFLE_SEED0 = 100


def configs(n=N):
    import workloads
    return {
        "cfg2": workloads.config2(n, seed=0),
        "cfg3": workloads.config3(n, hidden=(21, 21), seed=1),
        "cfg4": workloads.config4(n, seed=2),
    }


def disp(r):
    return r["last_center_x"] - r["first_center_x"]


def main():
    from oracle import oracle
    threads = os.cpu_count() or 1
    out = {}
    for name, (loco, robots) in configs().items():
        bd = loco.compile(robots)
        t0 = time.time()
        canon, _ = oracle.run(bd, n_threads=threads)
        assert np.all(canon["status"] == 0)
        shuf = [disp(oracle.run(bd, order=oracle.ORDER_SHUFFLED, seed=SHUFFLE_SEED0 + k, n_threads=threads)[0])
                for k in range(K + 1)]
        out[name + "_canon"] = disp(canon)
        # the same robots under ContinuousDetectionMode.NONE (canonical order): what the time-of-impact pass does to a
        # population is read off against the reorder null (tests/test_oracle_cpu.py)
        import importlib
        H = importlib.import_module("2dhmsr_b200")
        st = H.Settings()
        for f, _t in st.raw()._fields_:
            setattr(st.raw(), f, getattr(loco.settings.raw(), f))
        nopass, _ = oracle.run(H.Locomotion(20.0, loco.ground_profile, st.set(continuous_detection=0)).compile(robots),
                               n_threads=threads)
        out[name + "_canon_nopass"] = disp(nopass)
        out[name + "_time"] = canon["t_last"] - canon["t_first"]
        out[name + "_shuf"] = np.stack(shuf)
        print(f"{name}: {time.time() - t0:.0f} s, mean d {disp(canon).mean():.3f}, median sigma_perm "
              f"{np.median(np.std(np.stack(shuf[:K]), axis=0, ddof=1)):.3f}", flush=True)
    src = open(os.path.join(ROOT, "oracle", "vsr_oracle.c"), "rb").read()
    out["meta"] = np.array(json.dumps({"n": N, "k": K, "shuffle_seed0": SHUFFLE_SEED0,
                                       "oracle_sha256": hashlib.sha256(src).hexdigest()}))
    np.savez_compressed(os.path.join(HERE, "fp32_protocol.npz"), **out)


if __name__ == "__main__":
    main()
