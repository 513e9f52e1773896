"""Regenerates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

Three kinds of vectors, kept apart on purpose:
  reference_*   values that come from the REFERENCE itself (its tests, its source comments,
                its shipped asset assets/frames.png) -- typed in by hand below, never computed;
  survey_*      known answers recorded in SURVEY.md section 8(c) (java.util.Random port);
  oracle_*      outputs of OUR oracle, committed only as a regression pin of the oracle
                (they are NOT reference vectors: the dyn4j boundary is parity-unpinned).
"""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
H = importlib.import_module("2dhmsr_b200")
from oracle import oracle  # noqa: E402
import helpers  # noqa: E402

gold = {
    # MultiLayerPerceptronTest.java:50-95
    "reference_mlp_apply": {"activation": "RELU", "neurons": [1, 2, 1], "weights": [1, 0, 1, 2, 1, -1, 1],
                            "input": [2], "output": [5]},
    "reference_mlp_flat": {"neurons": [2, 2, 1], "flat": [1, 2, 3, 4, 5, 6, 7, 8, 9],
                           "unflat": [[[1, 2, 3], [4, 5, 6]], [[7, 8, 9]]]},
    # assets/frames.png = output of Example.main (Example.java:40-62): InfoDrawer prints the mean of
    # the VoxelPoly centres (InfoDrawer.java:73-83) with %5.1f at the 5 rendered frames
    # (FramesImageBuilder.java:85-92: first tick with t >= 4, then every tick with t - lastT >= 0.2).
    "reference_frames_png": {"labels": ["4,0", "4,2", "4,4", "4,7", "4,9"],
                             "centres": [[19.8, 11.0], [19.9, 10.3], [20.6, 11.0], [21.6, 11.6], [22.6, 11.8]],
                             "resolution": 0.05},
    # Locomotion.java:79 comment "the 1st double of nextDouble() is always around 0.73"
    "reference_random_first_double_about": 0.73,
    # SURVEY.md 8(c)(2)
    "survey_java_random": {"seed0_nextDouble": 0.730967787376657, "seed0_nextInt": -1155484576,
                           "seed42_nextInt": -1170105035, "seed0_nextGaussian": 0.8025330637390305},
    "survey_hilly_1_10_0": {"n_points": 205, "interior": [[22.0063, -0.9015], [37.2086, -0.1378], [49.6701, -1.8212]]},
    "survey_ticks": {"final_t": 20, "n": 1200, "last": 20.000000000000146},
}

# full hilly terrain from our JavaRandom port (pinned by the three interior points above)
t = H.Locomotion.create_terrain("hilly-1-10-0")
gold["port_hilly_1_10_0"] = {"xs": t[0], "ys": t[1]}

# oracle regression pins
loco = helpers.locomotion()
bd = loco.compile([helpers.config1_robot()])
res, pr = oracle.run(bd, trajectory=True)
cx, cy = helpers.poly_centres(pr["trajectory"][0])
gold["oracle_config1"] = {
    "note": "oracle output (canonical order, reduced spring mass), regression pin only",
    "result": {k: float(res[0][k]) for k in res.dtype.names},
    "centre_at_frames": [[float(cx[k]), float(cy[k])] for k in (240, 253, 266, 279, 292)],
}
bd = helpers.locomotion(5.0).compile(helpers.phase_robots(3, seed=7))
res, _ = oracle.run(bd)
gold["oracle_phase3"] = {"distance": [float(r["last_center_x"] - r["first_center_x"]) for r in res]}
bd = helpers.locomotion(3.0).compile(helpers.mlp_robots(2, hidden=(5,), seed=3))
res, _ = oracle.run(bd)
gold["oracle_mlp2"] = {"distance": [float(r["last_center_x"] - r["first_center_x"]) for r in res]}

with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(gold, f, indent=1)
print("wrote", os.path.join(HERE, "golden.json"))
