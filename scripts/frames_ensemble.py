"""Config 1 (Example.java) under every oracle switch combination vs the centres printed on assets/frames.png.

Prints the residual table of DESIGN.md section 2: rows = (spring mass, intra-robot contacts, CCD, order), columns = the five
frames; entries = |centre - reference| (max over x, y), plus the spread of a seeded-shuffle ensemble."""
import sys, os, json, itertools
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import helpers
from oracle import oracle
GOLD = json.load(open(os.path.join(R, "tests", "golden", "golden.json")))
g = np.asarray(GOLD["reference_frames_png"]["centres"])
TICKS = (240, 253, 266, 279, 292)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
names = {0: "canonical", 1: "insertion", 3: "island DFS"}
print("| spring mass | mass-mass contacts | CCD | order | centre at t=4.02 | centre at t=4.88 | max residual over the 5 frames |")
print("|---|---|---|---|---|---|---|")
for sm, sc, ccd in itertools.product((0, 1), (1, 0), (0, 1)):
    bd = helpers.locomotion(5.0, spring_mass_mode=sm, self_collision=sc).compile([helpers.config1_robot()])
    def run(order, seed=0):
        _, pr = oracle.run(bd, order=order, seed=seed, trajectory=True, self_collision=bool(sc), ccd=bool(ccd))
        cx, cy = helpers.poly_centres(pr["trajectory"][0])
        return np.asarray([[cx[k], cy[k]] for k in TICKS])
    for order in (0, 1, 3):
        c = run(order)
        print(f"| {'reduced' if sm == 0 else 'effective'} | {'on' if sc else 'off'} | {'on' if ccd else 'off'} | {names[order]} | "
              f"({c[0,0]:.1f}, {c[0,1]:.1f}) | ({c[4,0]:.1f}, {c[4,1]:.1f}) | {np.abs(c - g).max():.2f} |")
    ens = np.asarray([run(2, s) for s in range(K)])
    m, sd = ens.mean(0), ens.std(0)
    print(f"| {'reduced' if sm == 0 else 'effective'} | {'on' if sc else 'off'} | {'on' if ccd else 'off'} | {K} shuffles: mean +- std | "
          f"({m[0,0]:.1f}+-{sd[0,0]:.1f}, {m[0,1]:.1f}+-{sd[0,1]:.1f}) | ({m[4,0]:.1f}+-{sd[4,0]:.1f}, {m[4,1]:.1f}+-{sd[4,1]:.1f}) | "
          f"best member {np.abs(ens - g).max((1, 2)).min():.2f} |")
print("| reference (frames.png) | | | | (19.8, 11.0) | (22.6, 11.8) | 0 |")
