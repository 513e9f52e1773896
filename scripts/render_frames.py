"""Render the oracle's config-1 robot at the frames.png ticks, same framing guess (PIL)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from PIL import Image, ImageDraw
import helpers
from oracle import oracle
sm = int(sys.argv[1]); sc = int(sys.argv[2]); order = int(sys.argv[3]); out = sys.argv[4]
ticks = [int(a) for a in sys.argv[5].split(",")] if len(sys.argv) > 5 else [240, 253, 266, 279, 292]
bd = helpers.locomotion(6.0, spring_mass_mode=sm, self_collision=sc).compile([helpers.config1_robot()])
_, pr = oracle.run(bd, order=order, trajectory=True, self_collision=bool(sc))
tr = pr["trajectory"][0]
cx, cy = helpers.poly_centres(tr)
W, Hh, S = 300, 200, 11.7
im = Image.new("RGB", (W * len(ticks), Hh), "white"); d = ImageDraw.Draw(im)
h = 0.525
for f, k in enumerate(ticks):
    ox, oy = cx[k], cy[k]
    def P(x, y): return (f * W + W / 2 + (x - ox) * S, Hh / 2 - (y - oy) * S - 10)
    d.line([P(ox - 20, 5), P(ox + 20, 5)], fill="gray")
    nb = tr.shape[1]
    for b in range(nb):
        x, y, th = tr[k, b, :3]
        c, s = np.cos(th), np.sin(th)
        pts = [P(x + c * lx - s * ly, y + s * lx + c * ly) for lx, ly in ((-h, -h), (h, -h), (h, h), (-h, h))]
        d.polygon(pts, outline="black")
    for v in range(nb // 4):
        lx = [-h, h, h, -h]; ly = [h, h, -h, -h]
        pts = []
        for q in range(4):
            x, y, th = tr[k, v * 4 + q, :3]; c, s = np.cos(th), np.sin(th)
            pts.append(P(x + c * lx[q] - s * ly[q], y + s * lx[q] + c * ly[q]))
        d.polygon(pts, outline="blue")
    d.text((f * W + 4, 2), "t=%.2f pos=(%.1f,%.1f)" % ((k + 1) / 60, ox, oy), fill="blue")
im.save(out)
