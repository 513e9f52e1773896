"""The per-tick view of an episode: ``Outcome.Observation`` / ``VoxelPoly`` built from the device's
per-tick dumps, the rest of ``tasks/locomotion/Outcome.java`` on top of them, and
``behavior/BehaviorUtils.java`` (postures, footprints, gaits, spectra).

The device records, for the first ``trajectory_robots`` robots of a batch, the bodies
(``vsr_batch_trajectory``) and per voxel whether it touches the ground and the applied force
(``vsr_batch_voxel_trace``); that is everything ``Voxel.getVoxelPoly`` (Voxel.java:605-616) reads.
Everything here is host-side arithmetic in double, as in the reference.
"""
from __future__ import annotations

import math
from bisect import bisect_left
from collections import Counter
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

Point2 = Tuple[float, float]


class BoundingBox:
    """util/BoundingBox.java"""

    def __init__(self, min_: Point2, max_: Point2):
        self.min, self.max = min_, max_

    @staticmethod
    def largest(boxes: Iterable["BoundingBox"]) -> "BoundingBox":
        bs = list(boxes)
        return BoundingBox((min(b.min[0] for b in bs), min(b.min[1] for b in bs)),
                           (max(b.max[0] for b in bs), max(b.max[1] for b in bs)))


class VoxelPoly:
    """core/snapshots/VoxelPoly.java: the 4 outer vertexes of a voxel and its scalar state."""

    __slots__ = ("_v", "angle", "linear_velocity", "touching", "area_ratio", "area_ratio_energy",
                 "last_applied_force", "control_energy")

    def __init__(self, vertexes, angle, linear_velocity, touching, area_ratio, area_ratio_energy,
                 last_applied_force=0.0, control_energy=0.0):
        self._v = [(float(x), float(y)) for x, y in vertexes]
        self.angle = float(angle)
        self.linear_velocity = (float(linear_velocity[0]), float(linear_velocity[1]))
        self.touching = bool(touching)
        self.area_ratio = float(area_ratio)
        self.area_ratio_energy = float(area_ratio_energy)
        self.last_applied_force = float(last_applied_force)
        self.control_energy = float(control_energy)

    def vertexes(self) -> List[Point2]:
        return list(self._v)

    def bounding_box(self) -> BoundingBox:  # Poly.java:51-64
        xs, ys = [p[0] for p in self._v], [p[1] for p in self._v]
        return BoundingBox((min(xs), min(ys)), (max(xs), max(ys)))

    def area(self) -> float:  # Poly.java:40-49
        a, v, n = 0.0, self._v, len(self._v)
        for i in range(n):
            a = a + v[i][0] * (v[(n + i + 1) % n][1] - v[(n + i - 1) % n][1])
        return 0.5 * abs(a)

    def center(self) -> Point2:  # Poly.java:35-38 = Point2.average
        n = len(self._v)
        return (sum(p[0] for p in self._v) / n, sum(p[1] for p in self._v) / n)

    def get_angle(self) -> float:
        return self.angle

    def get_linear_velocity(self) -> Point2:
        return self.linear_velocity

    def is_touching_ground(self) -> bool:
        return self.touching

    def get_area_ratio(self) -> float:
        return self.area_ratio

    def get_area_ratio_energy(self) -> float:
        return self.area_ratio_energy

    def get_last_applied_force(self) -> float:
        return self.last_applied_force

    def get_control_energy(self) -> float:
        return self.control_energy


class Observation:
    """Outcome.java:38-39: record Observation(Grid<VoxelPoly> voxelPolies, terrainHeight, computationTime)."""

    __slots__ = ("voxel_polies", "terrain_height", "computation_time")

    def __init__(self, voxel_polies, terrain_height: float, computation_time: float):
        self.voxel_polies = voxel_polies
        self.terrain_height = float(terrain_height)
        self.computation_time = float(computation_time)


class Footprint:
    """behavior/Footprint.java"""

    def __init__(self, mask: Sequence[bool]):
        self.mask = tuple(bool(b) for b in mask)

    def get_mask(self):
        return self.mask

    def length(self) -> int:
        return len(self.mask)

    def __eq__(self, o):
        return isinstance(o, Footprint) and self.mask == o.mask

    def __hash__(self):
        return hash(self.mask)

    def __repr__(self):
        return "".join("_" if b else "." for b in self.mask)


class Gait:
    """behavior/Gait.java"""

    def __init__(self, footprints, mode_interval, coverage, duration, purity):
        self.footprints, self.mode_interval, self.coverage = list(footprints), mode_interval, coverage
        self.duration, self.purity = duration, purity

    def get_avg_touch_area(self) -> float:
        v = [sum(1.0 if b else 0.0 for b in f.mask) / f.length() for f in self.footprints]
        return sum(v) / len(v) if v else 0.0

    def get_coverage(self):
        return self.coverage

    def get_duration(self):
        return self.duration

    def get_footprints(self):
        return self.footprints

    def get_mode_interval(self):
        return self.mode_interval

    def get_purity(self):
        return self.purity

    def __repr__(self):
        return "Gait{footprints=%s, modeInterval=%.1fs, coverage=%.2f, duration=%.1fs, purity=%.2f}" % (
            self.footprints, self.mode_interval, self.coverage, self.duration, self.purity)


def _jdiv(a: float, b: float) -> float:
    """Java double division: x / 0 is +-Infinity or NaN, never an exception."""
    if b != 0.0:
        return a / b
    return math.nan if a == 0.0 or a != a else math.copysign(math.inf, a)


def _java_round(x: float) -> int:
    """Math.round: floor(x + 0.5)."""
    return int(math.floor(x + 0.5))


class BehaviorUtils:
    """behavior/BehaviorUtils.java"""

    @staticmethod
    def center(shapes) -> Point2:  # :47-56
        x = y = 0.0
        shapes = list(shapes)
        for s in shapes:
            c = s.center()
            x = x + c[0]
            y = y + c[1]
        return (x / len(shapes), y / len(shapes))

    @staticmethod
    def compute_average_posture(postures):  # :58-66
        from .hmsrobots import Grid
        postures = list(postures)
        w, h = postures[0].w, postures[0].h
        n = float(len(postures))
        return Grid.create(w, h, lambda x, y: sum(1.0 if p.get(x, y) else 0.0 for p in postures) / n > 0.5)

    @staticmethod
    def compute_footprint(polies, n: int) -> Footprint:  # :68-91
        polies = list(polies)
        if not polies:
            raise ValueError("Empty robot")
        boxes = [p.bounding_box() for p in polies]
        mn, mx = min(b.min[0] for b in boxes), max(b.max[0] for b in boxes)
        mask = [False] * n
        for p, b in zip(polies, boxes):
            if not p.is_touching_ground():
                continue
            lo = _java_round((b.min[0] - mn) / (mx - mn) * (n - 1))
            hi = _java_round((b.max[0] - mn) / (mx - mn) * (n - 1))
            for x in range(lo, min(hi, n - 1) + 1):
                mask[x] = True
        return Footprint(mask)

    @staticmethod
    def compute_posture(shapes, n: int):  # :177-219
        from .hmsrobots import Grid
        boxes = [s.bounding_box() for s in shapes]
        if not boxes:
            raise ValueError("Empty robot")
        mnx, mxx = min(b.min[0] for b in boxes), max(b.max[0] for b in boxes)
        mny, mxy = min(b.min[1] for b in boxes), max(b.max[1] for b in boxes)
        if (mxy - mny) < (mxx - mnx):
            d = (mxx - mnx) - (mxy - mny)
            mxy, mny = mxy + d / 2, mny - d / 2
        elif (mxy - mny) > (mxx - mnx):
            d = (mxy - mny) - (mxx - mnx)
            mxx, mnx = mxx + d / 2, mnx - d / 2
        mask = Grid.create(n, n, False)
        for b in boxes:
            x0 = _java_round((b.min[0] - mnx) / (mxx - mnx) * (n - 1))
            x1 = _java_round((b.max[0] - mnx) / (mxx - mnx) * (n - 1))
            y0 = _java_round((b.min[1] - mny) / (mxy - mny) * (n - 1))
            y1 = _java_round((b.max[1] - mny) / (mxy - mny) * (n - 1))
            for x in range(x0, x1 + 1):
                for y in range(y0, y1 + 1):
                    mask.set(x, y, True)
        return mask

    @staticmethod
    def compute_quantized_footprints(interval: float, polies: "Dict[float, list]", n: int) -> "Dict[float, Footprint]":  # :221-246
        keys = sorted(polies)
        out: Dict[float, Footprint] = {}
        t = keys[0]
        while t <= keys[-1]:
            lo, hi = bisect_left(keys, t), bisect_left(keys, t + interval)  # subMap(t, t + interval)
            local = [BehaviorUtils.compute_footprint(polies[k], n) for k in keys[lo:hi]]
            tot = float(len(local))
            counts = [sum(1.0 if f.mask[x] else 0.0 for f in local) for x in range(n)]
            out[t] = Footprint([c > tot / 2.0 for c in counts])
            t = t + interval
        return out

    @staticmethod
    def _mode(values):
        cnt = Counter(values)
        best = max(cnt.values())
        for v in values:  # first maximal entry in encounter order
            if cnt[v] == best:
                return v
        raise ValueError("empty")

    @staticmethod
    def compute_gaits(footprints: "Dict[float, Footprint]", min_len: int, max_len: int, interval: float) -> List[Gait]:  # :93-155
        keys = sorted(footprints)
        fl = [footprints[k] for k in keys]
        ranges = [(k, k + interval) for k in keys]
        seqs: Dict[tuple, list] = {}
        for l in range(min_len, max_len + 1):
            for i in range(l, len(fl) + 1):
                seq = tuple(fl[i - l:i])
                local = seqs.get(seq, [])
                if not local or local[-1][1] <= ranges[i - l][0]:
                    local.append((ranges[i - l][0], ranges[i - 1][1]))
                seqs[seq] = local
        all_int = [l[i + 1][0] - l[i][1] for l in seqs.values() for i in range(len(l) - 1)]
        if not all_int:
            return []
        mode_int = BehaviorUtils._mode(all_int)
        out = []
        for seq, l in seqs.items():
            if len(l) <= 1:
                continue
            ints = [l[i + 1][0] - l[i][1] for i in range(len(l) - 1)]
            cov = [_jdiv(l[i][1] - l[i][0], ints[i]) for i in range(len(ints))]
            lm = BehaviorUtils._mode(ints)
            g = Gait(seq, lm, sum(cov) / len(cov) if cov else 0.0, sum(b - a for a, b in l),
                     sum(1 for d in ints if d == lm) / float(len(l)))
            if g.mode_interval == mode_int:
                out.append(g)
        return out

    @staticmethod
    def compute_main_gait(interval: float, longest_interval: float, polies: "Dict[float, list]", n: int) -> Optional[Gait]:  # :157-175
        gaits = BehaviorUtils.compute_gaits(BehaviorUtils.compute_quantized_footprints(interval, polies, n), 2,
                                            _java_round(longest_interval / interval), interval)
        if not gaits:
            return None
        gaits.sort(key=lambda g: -g.duration)
        return gaits[0]

    @staticmethod
    def compute_spectrum(signal, dt: Optional[float] = None) -> "Dict[float, float]":  # :262-287
        """``signal``: {t: value} (dT = mean key spacing) or a sequence with an explicit ``dt``."""
        if isinstance(signal, dict):
            keys = sorted(signal)
            iv = [keys[i + 1] - keys[i] for i in range(len(keys) - 1)]
            dt = sum(iv) / len(iv) if iv else 0.0
            vals = np.array([signal[k] for k in keys], dtype=np.float64)
        else:
            vals = np.asarray(signal, dtype=np.float64)
        padded = int(math.pow(2.0, math.ceil(math.log(len(vals)) / math.log(2.0))))
        if padded != len(vals):
            vals = np.concatenate([vals, np.zeros(padded - len(vals))])
        f = np.abs(np.fft.fft(vals))[: padded // 2 + 1]  # DftNormalization.STANDARD = unnormalised forward
        return {1.0 / dt / 2.0 * i / float(len(f)): float(f[i]) for i in range(len(f))}

    @staticmethod
    def compute_quantized_spectrum(signal, min_f: float, max_f: float, n_bins: int) -> "Dict[Tuple[float, float], float]":  # :248-260
        spec = BehaviorUtils.compute_spectrum(signal)
        keys = sorted(spec)
        span = (max_f - min_f) / float(n_bins)
        out = {}
        for i in range(n_bins):
            lo, hi = min_f + span * i, min_f + span * (i + 1.0)
            sel = [spec[k] for k in keys[bisect_left(keys, lo):bisect_left(keys, hi)]]
            out[(lo, hi)] = sum(sel) / len(sel) if sel else 0.0
        return out

    @staticmethod
    def get_central_element(grid):  # :289-306
        cells = [(x, y) for (x, y), v in grid if v is not None]
        if not cells:
            raise ValueError("Cannot get central element of an empty grid")
        mx = sum(c[0] for c in cells) / len(cells)
        my = sum(c[1] for c in cells) / len(cells)
        best, bx, by = float("inf"), 0, 0
        for x in range(grid.w):
            for y in range(grid.h):
                d = (x - mx) * (x - mx) + (y - my) * (y - my)
                if d < best:
                    best, bx, by = d, x, y
        return grid.get(bx, by)


def ground_y_at(xs: Sequence[float], ys: Sequence[float], x: float) -> float:
    """Ground.yAt (Ground.java:104-111)."""
    for i in range(1, len(xs)):
        if xs[i - 1] <= x <= xs[i]:
            return (x - xs[i - 1]) * (ys[i] - ys[i - 1]) / (xs[i] - xs[i - 1]) + ys[i - 1]
    return float("nan")


def build_observations(body_grid, trajectory: np.ndarray, trace: np.ndarray, tick_t: np.ndarray, terrain,
                       side_length: float, mass_side: float, wall_time: float = 0.0) -> "Dict[float, Observation]":
    """The ``observations`` map of Locomotion.apply (Locomotion.java:193-203) for one robot.

    ``trajectory`` [ticks][4 * voxels][6] and ``trace`` [ticks][voxels][2] are the device's dumps (voxels
    in Grid.values() order, bodies NW, NE, SE, SW); the scalar getters follow Voxel.java:508-556."""
    from .hmsrobots import Grid
    nt = trajectory.shape[0]
    cells = [(x, y) for (x, y), v in body_grid if v is not None]
    nv = len(cells)
    tr = trajectory.reshape(nt, nv, 4, 6)
    h = 0.5 * mass_side
    lx = np.array([-1.0, 1.0, 1.0, -1.0]) * h   # getIndexedVertex(k, 3 - k): the outer corner of body k
    ly = np.array([1.0, 1.0, -1.0, -1.0]) * h
    c, s = np.cos(tr[..., 2]), np.sin(tr[..., 2])
    vx = tr[..., 0] + c * lx - s * ly            # [tick][voxel][4]
    vy = tr[..., 1] + s * lx + c * ly
    area = np.zeros((nt, nv))
    for i in range(4):
        area = area + vx[..., i] * (vy[..., (i + 1) % 4] - vy[..., (i + 3) % 4])
    ratio = 0.5 * np.abs(area) / side_length / side_length
    up = np.arctan2(tr[:, :, 1, 1] - tr[:, :, 0, 1], tr[:, :, 1, 0] - tr[:, :, 0, 0])
    down = np.arctan2(tr[:, :, 2, 1] - tr[:, :, 3, 1], tr[:, :, 2, 0] - tr[:, :, 3, 0])
    angle = (up + down) / 2.0
    lvx = (((tr[:, :, 0, 3] + tr[:, :, 1, 3]) + tr[:, :, 2, 3]) + tr[:, :, 3, 3]) / 4.0
    lvy = (((tr[:, :, 0, 4] + tr[:, :, 1, 4]) + tr[:, :, 2, 4]) + tr[:, :, 3, 4]) / 4.0
    force = trace[:, :, 1]
    # Voxel.act (Voxel.java:197-203) runs before the controller: the energy of tick k holds the force
    # applied at tick k - 1
    are = np.cumsum(ratio * ratio, axis=0)
    prev = np.concatenate([np.zeros((1, nv)), force[:-1]], axis=0)
    ce = np.cumsum(prev * prev, axis=0)
    xs, ys = terrain
    out: Dict[float, Observation] = {}
    for k in range(nt):
        polies = Grid.create(body_grid.w, body_grid.h, None)
        for j, (gx, gy) in enumerate(cells):
            polies.set(gx, gy, VoxelPoly(zip(vx[k, j], vy[k, j]), angle[k, j], (lvx[k, j], lvy[k, j]),
                                         trace[k, j, 0] != 0.0, ratio[k, j], are[k, j], force[k, j], ce[k, j]))
        # Robot.center (Robot.java:122-125) = average of the voxels' centres = average of their 4 body centres
        cx = float(np.mean(tr[k, :, :, 0].mean(axis=1)))
        out[float(tick_t[k])] = Observation(polies, ground_y_at(xs, ys, cx), wall_time * (k + 1) / nt)
    return out
