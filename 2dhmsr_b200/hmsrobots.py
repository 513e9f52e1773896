"""Host-side mirror of the reference's Java API for the Locomotion hot path.

Same names, argument meaning and error behaviour as
``it.units.erallab.hmsrobots`` (file:line cited per item), restricted to what the
hot path needs: describing robots / terrain / controllers and handing a *batch*
of them to the CUDA stepper through the C ABI (``include/vsr_b200.h``).

Nothing here steps physics: ``Locomotion.apply`` compiles a ``vsr_batch_desc``
and calls ``libvsr_b200.so``.  ``IllegalArgumentException`` -> ``ValueError``.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from dataclasses import dataclass, field
from typing import Callable, Generic, Iterable, Iterator, List, Optional, Sequence, Tuple, TypeVar

import numpy as np

from . import _abi as A

T = TypeVar("T")


# --------------------------------------------------------------------------- java.util.Random
class JavaRandom:
    """Bit-exact ``java.util.Random`` (48-bit LCG; polar-method nextGaussian).

    Needed by Locomotion.createTerrain (Locomotion.java:76-117).  Java uses
    StrictMath.log/sqrt in nextGaussian; ``math.log``/``math.sqrt`` (glibc) agree
    to <= 1 ulp.
    """

    _MULT = 0x5DEECE66D
    _MASK = (1 << 48) - 1

    def __init__(self, seed: int):
        self._seed = (int(seed) ^ self._MULT) & self._MASK
        self._next_gaussian: Optional[float] = None

    def _next(self, bits: int) -> int:
        self._seed = (self._seed * self._MULT + 0xB) & self._MASK
        v = self._seed >> (48 - bits)
        if v >= 1 << (bits - 1) and bits == 32:
            v -= 1 << 32
        return v

    def next_int(self, bound: Optional[int] = None) -> int:
        if bound is None:
            return self._next(32)
        if bound <= 0:
            raise ValueError("bound must be positive")
        if (bound & -bound) == bound:
            return (bound * self._next(31)) >> 31
        while True:
            bits = self._next(31)
            val = bits % bound
            if bits - val + (bound - 1) < (1 << 31):
                return val

    def next_double(self) -> float:
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))

    def next_gaussian(self) -> float:
        if self._next_gaussian is not None:
            g, self._next_gaussian = self._next_gaussian, None
            return g
        while True:
            v1 = 2 * self.next_double() - 1
            v2 = 2 * self.next_double() - 1
            s = v1 * v1 + v2 * v2
            if 0 < s < 1:
                break
        mul = math.sqrt(-2 * math.log(s) / s)
        self._next_gaussian = v2 * mul
        return v1 * mul


# --------------------------------------------------------------------------- util.Grid
class Grid(Generic[T]):
    """util/Grid.java: storage ``ts[y*w+x]`` (:180-188); ``create(w,h,filler)`` calls
    the filler x-outer / y-inner (:103-111) while iteration and ``values()`` are
    row-major, y-outer (:69-75,268-270)."""

    def __init__(self, w: int, h: int):
        self.w, self.h = w, h
        self._ts: List[Optional[T]] = [None] * (w * h)

    @staticmethod
    def create(w: int, h: int, filler=None) -> "Grid":
        g = Grid(w, h)
        for x in range(w):
            for y in range(h):
                g.set(x, y, filler(x, y) if callable(filler) else filler)
        return g

    @staticmethod
    def create_from(other: "Grid", fn: Callable) -> "Grid":
        g = Grid(other.w, other.h)
        for (x, y), v in other:
            g.set(x, y, fn(v))
        return g

    def get(self, x: int, y: int):
        if x < 0 or x >= self.w or y < 0 or y >= self.h:
            return None
        return self._ts[y * self.w + x]

    def set(self, x: int, y: int, v) -> None:
        self._ts[y * self.w + x] = v

    def values(self) -> list:
        return list(self._ts)

    def __iter__(self) -> Iterator[Tuple[Tuple[int, int], Optional[T]]]:
        for c in range(self.w * self.h):
            y, x = divmod(c, self.w)
            yield (x, y), self._ts[c]

    def count(self, pred) -> int:
        return sum(1 for v in self._ts if pred(v))


@dataclass(frozen=True)
class DoubleRange:
    """util/DoubleRange.java:29-58."""
    min: float
    max: float

    def __post_init__(self):
        if self.max < self.min:
            raise ValueError(f"Max has to be lower or equal than min; {self.max} is not than {self.min}.")

    def extent(self) -> float:
        return self.max - self.min

    def clip(self, v: float) -> float:
        return min(max(v, self.min), self.max)

    def normalize(self, v: float) -> float:
        return (self.clip(v) - self.min) / (self.max - self.min)


# --------------------------------------------------------------------------- sensors
class Sensor:
    """core/sensors/Sensor.java:26-34 (description only: readout runs on the device)."""

    def domains(self) -> List[DoubleRange]:
        raise NotImplementedError

    def encode(self) -> A.vsr_sensor:
        raise NotImplementedError


class Touch(Sensor):  # sensors/Touch.java
    def domains(self):
        return [DoubleRange(0.0, 1.0)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_TOUCH
        return s


class AreaRatio(Sensor):  # sensors/AreaRatio.java:21-37
    def domains(self):
        return [DoubleRange(0.5, 1.5)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_AREA_RATIO
        return s


class Velocity(Sensor):  # sensors/Velocity.java:29-80
    X, Y = "X", "Y"

    def __init__(self, rotated: bool, max_velocity_norm: float, *axes: str):
        self.rotated, self.max_velocity_norm = rotated, max_velocity_norm
        self.axes = [a for a in (self.X, self.Y) if a in axes]  # EnumSet order

    def domains(self):
        return [DoubleRange(-self.max_velocity_norm, self.max_velocity_norm) for _ in self.axes]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_VELOCITY
        s.axes = (A.VSR_AXIS_X if self.X in self.axes else 0) | (A.VSR_AXIS_Y if self.Y in self.axes else 0)
        s.rotated = 1 if self.rotated else 0
        s.max_velocity_norm = self.max_velocity_norm
        return s


class Angle(Sensor):  # sensors/Angle.java
    def domains(self):
        return [DoubleRange(-math.pi, math.pi)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_ANGLE
        return s


class Constant(Sensor):  # sensors/Constant.java (one value per sensor on the accelerated path)
    def __init__(self, *values: float):
        if len(values) != 1:
            raise NotImplementedError("Constant: one value per sensor on the accelerated path")
        self.values = [float(v) for v in values]

    def domains(self):
        lo, hi = min(0.0, min(self.values)), max(1.0, max(self.values))
        return [DoubleRange(lo, hi) for _ in self.values]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_CONSTANT
        s.max_velocity_norm = self.values[0]
        return s


class AppliedForce(Sensor):  # sensors/AppliedForce.java
    def domains(self):
        return [DoubleRange(-1.0, 1.0)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_APPLIED_FORCE
        return s


class Malfunction(Sensor):  # sensors/Malfunction.java (a non-breakable voxel is never broken)
    def domains(self):
        return [DoubleRange(0.0, 1.0)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_MALFUNCTION
        return s


class Crumpling(Sensor):  # sensors/Crumpling.java
    def domains(self):
        return [DoubleRange(0.0, 1.0)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_CRUMPLING
        return s


class ControlPower(Sensor):  # sensors/ControlPower.java
    def __init__(self, control_interval: float):
        self.control_interval = float(control_interval)

    def domains(self):
        return [DoubleRange(0.0, 1.0 / self.control_interval)]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_CONTROL_POWER
        s.max_velocity_norm = self.control_interval
        return s


class Lidar(Sensor):
    """sensors/Lidar.java: rays from the voxel centre; a reading is the distance to the terrain (robot
    parts are filtered out, :37-46) or the ray length."""
    SIDES = {  # :66-86 (startAngle, endAngle)
        "N": (math.pi / 4.0, math.pi * 3.0 / 4.0), "E": (math.pi * -1.0 / 4.0, math.pi / 4.0),
        "S": (math.pi * 5.0 / 4.0, math.pi * 7.0 / 4.0), "W": (math.pi * 3.0 / 4.0, math.pi * 5.0 / 4.0),
    }

    def __init__(self, ray_length: float, *ray_directions: float):
        if not 1 <= len(ray_directions) <= A.VSR_LIDAR_MAX_RAYS:
            raise NotImplementedError(f"Lidar: 1..{A.VSR_LIDAR_MAX_RAYS} rays per sensor on the accelerated path")
        self.ray_length, self.ray_directions = float(ray_length), [float(d) for d in ray_directions]

    @staticmethod
    def sample_range_with_rays(n: int, start: float, end: float) -> List[float]:  # :88-94
        if n == 1:
            return [(end - start) / 2]
        out, d = [], start
        for _ in range(n):  # DoubleStream.iterate: running sum
            out.append(d)
            d = d + (end - start) / (n - 1)
        return out

    @staticmethod
    def from_sides(ray_length: float, rays_per_side: dict) -> "Lidar":  # :57-65
        dirs: List[float] = []
        for side, n in rays_per_side.items():
            for d in Lidar.sample_range_with_rays(n, *Lidar.SIDES[side]):
                if d not in dirs:  # .distinct()
                    dirs.append(d)
        return Lidar(ray_length, *dirs)

    def domains(self):
        return [DoubleRange(0.0, self.ray_length) for _ in self.ray_directions]

    def encode(self):
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_LIDAR
        s.axes = len(self.ray_directions)
        s.max_velocity_norm = self.ray_length
        for i, d in enumerate(self.ray_directions):
            s.lidar_dirs[i] = d
        return s


class TimeFunction(Sensor):
    """sensors/TimeFunction.java.  The accelerated path evaluates ``t -> sin(2 pi f t)`` on [-1, 1] (the
    "cpg" sensor of RobotUtils.java:57); build it with ``TimeFunction.sin(f)``."""

    def __init__(self, function, min_: float, max_: float, _freq: Optional[float] = None):
        self.function, self.min, self.max, self._freq = function, min_, max_, _freq

    @staticmethod
    def sin(frequency: float) -> "TimeFunction":
        return TimeFunction(lambda t: math.sin(2 * math.pi * frequency * t), -1.0, 1.0, frequency)

    def domains(self):
        return [DoubleRange(self.min, self.max)]

    def encode(self):
        if self._freq is None or (self.min, self.max) != (-1.0, 1.0):
            raise NotImplementedError("TimeFunction: only TimeFunction.sin(f) on [-1, 1] runs on the device")
        s = A.vsr_sensor()
        s.base = A.VSR_SENSOR_TIME_SIN
        s.max_velocity_norm = self._freq
        return s


class _Aggregator(Sensor):  # sensors/AggregatorSensor.java:31-66
    KIND = A.VSR_AGG_NONE

    def __init__(self, sensor: Sensor, interval: float):
        if isinstance(sensor, (_Aggregator, SoftNormalization)):
            raise NotImplementedError("nested aggregators are outside the hot path")
        self.sensor, self.interval = sensor, interval

    def encode(self):
        s = self.sensor.encode()
        s.aggregator = self.KIND
        s.interval = self.interval
        return s


class Average(_Aggregator):  # sensors/Average.java
    KIND = A.VSR_AGG_AVERAGE

    def domains(self):
        return self.sensor.domains()


class Trend(_Aggregator):  # sensors/Trend.java:28-35
    KIND = A.VSR_AGG_TREND

    def domains(self):
        return [DoubleRange(-d.extent() / self.interval, d.extent() / self.interval) for d in self.sensor.domains()]


class DynamicNormalization(_Aggregator):
    """sensors/DynamicNormalization.java:26-61: every reading rescaled by the minimum and maximum of the readings kept in
    the window, clamped to [0, 1].  While the window holds a single distinct value (always at the first tick) the
    reference divides 0 by 0: the reading is NaN there, and it is NaN here."""
    KIND = A.VSR_AGG_DYNAMIC_NORM

    def domains(self):
        return [DoubleRange(0.0, 1.0) for _ in self.sensor.domains()]


class SoftNormalization(Sensor):  # sensors/SoftNormalization.java
    def __init__(self, sensor: Sensor):
        if isinstance(sensor, SoftNormalization):
            raise NotImplementedError("nested SoftNormalization is outside the hot path")
        self.sensor = sensor

    def domains(self):
        return [DoubleRange(0.0, 1.0) for _ in self.sensor.domains()]

    def encode(self):
        s = self.sensor.encode()
        s.soft_norm = 1
        return s


class Normalization(Sensor):  # sensors/Normalization.java
    def __init__(self, sensor: Sensor):
        if isinstance(sensor, (SoftNormalization, Normalization)):
            raise NotImplementedError("nested normalizations are outside the hot path")
        self.sensor = sensor

    def domains(self):
        return [DoubleRange(0.0, 1.0) for _ in self.sensor.domains()]

    def encode(self):
        s = self.sensor.encode()
        s.soft_norm = A.VSR_NORM_LINEAR
        return s


class Noisy(Sensor):
    """sensors/Noisy.java: the wrapped sensor's readings + ``Random(seed).nextGaussian() * sigma * extent`` of the
    reading's domain (:49-62).  The generator belongs to the sensor instance; a robot is evaluated from a
    freshly constructed state (the stream starts at the seed for every episode)."""

    def __init__(self, sensor: Sensor, sigma: float, seed: int):
        if isinstance(sensor, Noisy):
            raise NotImplementedError("nested Noisy sensors are outside the hot path")
        self.sensor, self.sigma, self.seed = sensor, float(sigma), int(seed)

    def domains(self):
        return self.sensor.domains()

    def encode(self):
        s = self.sensor.encode()
        s.noise_sigma, s.noise_seed = self.sigma, self.seed
        return s


# --------------------------------------------------------------------------- Voxel / Robot
class Voxel:
    """core/objects/Voxel.java: the 13-argument constructor (:103-142) and the
    default one (:144-160).  Holds only the description."""

    SIDE_LENGTH = 3.0
    MASS_SIDE_LENGTH_RATIO = 0.35
    SPRING_F = 8.0
    SPRING_D = 0.3
    MASS_LINEAR_DAMPING = 0.1
    MASS_ANGULAR_DAMPING = 0.1
    FRICTION = 10.0
    RESTITUTION = 0.1
    MASS = 1.0

    def __init__(self, sensors: Sequence[Sensor], material: Optional[A.vsr_material] = None):
        self.sensors = list(sensors)
        self.material = material if material is not None else A.default_material()
        m = self.material
        lim = (2 * m.mass_side_length_ratio) ** 2
        if m.area_ratio_passive_min > -math.inf and m.area_ratio_passive_min < lim:  # Voxel.java:132-140
            raise ValueError(
                f"Min of areaRatioPassiveRange={m.area_ratio_passive_min} cannot be lower than the value "
                f"({lim}) imposed by massSideLengthRatio={m.mass_side_length_ratio}"
            )
        # Voxel.assemble builds the passive and active SpringRanges (Voxel.java:262-301); SpringRange (:189-193)
        # rejects min > rest, max < rest, min < 0 (comparisons with the NaN of an infinite passive bound are false)
        L, ms = m.side_length, m.side_length * m.mass_side_length_ratio
        for a_min, a_max in ((m.area_ratio_passive_min, m.area_ratio_passive_max), (m.area_ratio_active_min, m.area_ratio_active_max)):
            s_min = math.sqrt(L * L * a_min) if a_min >= 0 else math.nan
            s_max = math.sqrt(L * L * a_max) if a_max >= 0 else math.nan
            pp = (s_min - 2 * ms, L - 2 * ms, s_max - 2 * ms)
            for lo_, rest_, hi_ in (pp,
                                    tuple(math.sqrt(ms * ms + q * q) for q in pp),
                                    ((s_min - ms) * math.sqrt(2.0), (L - ms) * math.sqrt(2.0), (s_max - ms) * math.sqrt(2.0))):
                if lo_ > rest_ or hi_ < rest_ or lo_ < 0:
                    raise ValueError(f"Wrong spring range [{lo_}, {rest_}, {hi_}]")

    def get_sensors(self) -> List[Sensor]:
        return self.sensors


class BreakableVoxel(Voxel):
    """core/objects/BreakableVoxel.java: a voxel whose ACTUATOR, SENSORS or STRUCTURE may malfunction (ZERO, FROZEN,
    RANDOM; STRUCTURE: FROZEN only, the others throw in the reference at the first break) when a CONTROL, AREA or TIME
    counter trips its threshold, restored ``restore_time`` after the last break.  ``malfunctions`` maps a component name
    to the set of its malfunction types, ``trigger_thresholds`` a trigger name to its threshold (both iterated in enum
    order on the device, see include/vsr_b200.h)."""

    COMPONENTS = ("ACTUATOR", "SENSORS", "STRUCTURE")
    TRIGGERS = ("CONTROL", "AREA", "TIME")
    TYPES = ("NONE", "ZERO", "FROZEN", "RANDOM")

    def __init__(self, sensors: Sequence[Sensor], random_seed: int, malfunctions: dict, trigger_thresholds: dict,
                 restore_time: float, material: Optional[A.vsr_material] = None):
        super().__init__(sensors, material)
        for c, ts in malfunctions.items():
            if c not in self.COMPONENTS or any(t not in self.TYPES for t in ts) or not ts:
                raise ValueError(f"unknown malfunction {c}: {ts}")
        if any(t not in ("NONE", "FROZEN") for t in malfunctions.get("STRUCTURE", ())):
            raise ValueError("Unsupported structure malfunction type.")   # BreakableVoxel.java:258
        if any(t not in self.TRIGGERS for t in trigger_thresholds):
            raise ValueError(f"unknown malfunction trigger in {trigger_thresholds}")
        self.random_seed, self.malfunctions = int(random_seed), {c: frozenset(ts) for c, ts in malfunctions.items()}
        self.trigger_thresholds, self.restore_time = dict(trigger_thresholds), float(restore_time)

    def encode_breakable(self) -> tuple:
        """(malfunction bit masks per component, thresholds per trigger (NaN = absent), restore time, seed)"""
        mal = tuple(sum(1 << self.TYPES.index(t) for t in self.malfunctions.get(c, ())) for c in self.COMPONENTS)
        thr = tuple(float(self.trigger_thresholds.get(t, math.nan)) for t in self.TRIGGERS)
        return mal, thr, self.restore_time, self.random_seed


class Controller:
    """core/controllers/Controller.java:29 (description only)."""
    KIND = -1

    def params(self, voxels: Grid, tick_times: np.ndarray) -> np.ndarray:
        raise NotImplementedError


def _present(voxels: Grid) -> List[Tuple[int, int]]:
    return [(x, y) for (x, y), v in voxels if v is not None]


class TimeFunctions(Controller):
    """controllers/TimeFunctions.java:49-64: opaque per-voxel lambdas of t.  The host
    samples each lambda at the tick times (the only way to keep arbitrary lambdas)."""
    KIND = A.VSR_CTRL_TABULATED

    def __init__(self, functions: Grid):
        self.functions = functions

    def params(self, voxels, tick_times):
        cells = _present(voxels)
        out = np.empty((len(tick_times), len(cells)), dtype=np.float64)
        for j, (x, y) in enumerate(cells):
            f = self.functions.get(x, y)
            if f is None:
                raise ValueError(f"no time function for voxel ({x},{y})")
            for k, t in enumerate(tick_times):
                out[k, j] = f(float(t))
        return out.reshape(-1)


class PhaseSin(TimeFunctions):
    """controllers/PhaseSin.java:50-66: sin(2*pi*frequency*t + phase) * amplitude."""
    KIND = A.VSR_CTRL_PHASE_SIN

    def __init__(self, frequency: float, amplitude: float, phases: Grid):
        self.frequency, self.amplitude, self.phases = frequency, amplitude, phases

    def params(self, voxels, tick_times):
        cells = _present(voxels)
        ph = []
        for x, y in cells:
            p = self.phases.get(x, y)
            if p is None:
                raise ValueError(f"no phase for voxel ({x},{y})")
            ph.append(float(p))
        return np.asarray([self.frequency, self.amplitude] + ph, dtype=np.float64)


class MultiLayerPerceptron:
    """controllers/MultiLayerPerceptron.java (flat weights, bias first, :139-151)."""

    def __init__(self, activation: str, n_input: int, inner: Sequence[int], n_output: int,
                 weights: Optional[Sequence[float]] = None):
        if activation not in A.VSR_ACT:
            raise ValueError(f"unknown activation {activation}")
        self.activation = activation
        self.neurons = [n_input, *inner, n_output]  # countNeurons :97-104
        n = self.count_weights(self.neurons)
        w = np.zeros(n) if weights is None else np.asarray(weights, dtype=np.float64)
        if w.size != n:  # :56-62
            raise ValueError(f"Wrong number of weights: {n} expected, {w.size} found")
        self.weights = w

    @staticmethod
    def count_weights(neurons: Sequence[int]) -> int:  # :106-112
        return sum(neurons[i] * (neurons[i - 1] + 1) for i in range(1, len(neurons)))

    @staticmethod
    def unflat(flat: Sequence[float], neurons: Sequence[int]):  # :153-166
        out, c = [], 0
        for i in range(1, len(neurons)):
            layer = []
            for _ in range(neurons[i]):
                layer.append(list(flat[c:c + neurons[i - 1] + 1]))
                c += neurons[i - 1] + 1
            out.append(layer)
        return out

    @staticmethod
    def flat(unflat, neurons: Sequence[int]):  # :139-151
        return [w for i in range(1, len(neurons)) for j in range(neurons[i]) for w in unflat[i - 1][j]]

    def get_params(self):
        return self.weights

    _ACT = {
        "RELU": lambda x: 0.0 if x < 0 else x, "SIGMOID": lambda x: 1.0 / (1.0 + math.exp(-x)),
        "SIN": math.sin, "TANH": math.tanh, "SIGN": lambda x: (x > 0) - (x < 0) + 0.0, "IDENTITY": lambda x: x,
    }

    def apply(self, x: Sequence[float]):
        """MultiLayerPerceptron.apply (:168-189) on the host, in double (reference of the device's
        arithmetic in tests; the product path evaluates controllers in the kernel)."""
        if len(x) != self.neurons[0]:
            raise ValueError(f"Expected input length is {self.neurons[0]}: found {len(x)}")
        f = self._ACT[self.activation]
        w = self.unflat(self.weights, self.neurons)
        a = [f(float(v)) for v in x]
        for i in range(1, len(self.neurons)):
            b = []
            for j in range(self.neurons[i]):
                s = w[i - 1][j][0]
                for k in range(1, self.neurons[i - 1] + 1):
                    s = s + a[k - 1] * w[i - 1][j][k]
                b.append(f(s))
            a = b
        return a

    def get_input_dimension(self):
        return self.neurons[0]

    def get_output_dimension(self):
        return self.neurons[-1]


class CentralizedSensing(Controller):
    """controllers/CentralizedSensing.java:68-127."""
    KIND = A.VSR_CTRL_MLP

    def __init__(self, voxels: Grid, function: MultiLayerPerceptron):
        self.n_inputs = self.n_of_inputs(voxels)
        self.n_outputs = self.n_of_outputs(voxels)
        self.set_function(function)

    @staticmethod
    def n_of_inputs(voxels: Grid) -> int:
        return sum(len(s.domains()) for v in voxels.values() if v is not None for s in v.get_sensors())

    @staticmethod
    def n_of_outputs(voxels: Grid) -> int:
        return sum(1 for v in voxels.values() if v is not None)

    def set_function(self, function: MultiLayerPerceptron) -> None:
        if function.get_input_dimension() != self.n_inputs or function.get_output_dimension() != self.n_outputs:
            raise ValueError(  # :121-127
                "Wrong dimension of input or output in provided function: "
                f"R^{self.n_inputs}->R^{self.n_outputs} expected, "
                f"R^{function.get_input_dimension()}->R^{function.get_output_dimension()} found"
            )
        self.function = function

    def params(self, voxels, tick_times):
        return np.asarray(self.function.weights, dtype=np.float64)


class DistributedSensing(Controller):
    """controllers/DistributedSensing.java: one function per voxel, fed with the voxel's sensor readings
    and with ``signals`` values from each of its N, E, S, W neighbours' previous outputs; output 0 is the
    control signal, the other ``4 * signals`` are emitted towards the neighbours (:141-207).  The
    accelerated path takes one MultiLayerPerceptron per voxel (Starter.java:176-190), all with the same
    activation and hidden layer sizes."""
    KIND = A.VSR_CTRL_DISTRIBUTED
    DIRS = ((0, -1, 0), (1, 0, 1), (0, 1, 2), (-1, 0, 3))  # N, E, S, W: dx, dy, index (:83-101)
    ADJACENT = (2, 3, 0, 1)                                  # Dir.adjacent: N<->S, E<->W

    def __init__(self, voxels: Grid, signals: int):
        self.signals = int(signals)
        self.n_of_input_grid = Grid.create(voxels.w, voxels.h, lambda x, y: 0 if voxels.get(x, y) is None
                                           else self.n_of_inputs_of(voxels.get(x, y), self.signals))
        self.n_of_output_grid = Grid.create(voxels.w, voxels.h, lambda x, y: 0 if voxels.get(x, y) is None
                                            else self.n_of_outputs_of(voxels.get(x, y), self.signals))
        self.functions = Grid.create(voxels.w, voxels.h, None)
        self.reset()

    @staticmethod
    def n_of_inputs_of(voxel: "Voxel", signals: int) -> int:  # :141-143
        return signals * 4 + sum(len(s.domains()) for s in voxel.get_sensors())

    @staticmethod
    def n_of_outputs_of(voxel: "Voxel", signals: int) -> int:  # :145-147
        return 1 + signals * 4

    def n_of_inputs(self, x: int, y: int) -> int:
        return self.n_of_input_grid.get(x, y)

    def n_of_outputs(self, x: int, y: int) -> int:
        return self.n_of_output_grid.get(x, y)

    def get_functions(self) -> Grid:
        return self.functions

    def reset(self) -> None:  # :209-222
        w, h = self.n_of_input_grid.w, self.n_of_input_grid.h
        self.last_signals = Grid.create(w, h, lambda x, y: [0.0] * (4 * self.signals)
                                        if self.n_of_output_grid.get(x, y) else None)

    def _last_signals(self, x: int, y: int):  # getLastSignals :188-207
        vals = [0.0] * (4 * self.signals)
        c = 0
        for dx, dy, _ in self.DIRS:
            ax, ay = x + dx, y + dy
            adj = self.last_signals.get(ax, ay) if 0 <= ax < self.last_signals.w and 0 <= ay < self.last_signals.h else None
            if adj is not None:
                i = self.ADJACENT[self.DIRS.index((dx, dy, _))]
                vals[c:c + self.signals] = adj[i * self.signals:(i + 1) * self.signals]
            c += self.signals
        return vals

    def compute_control_signals(self, t: float, readings: Grid) -> Grid:
        """computeControlSignals (:150-186) on a grid of per-voxel sensor readings (host-side reference
        of the device's arithmetic; the functions are evaluated with ``MultiLayerPerceptron.apply``)."""
        out = Grid.create(readings.w, readings.h, None)
        cur = Grid.create(readings.w, readings.h, None)
        for (x, y), r in readings:
            if r is None:
                continue
            inputs = list(r) + self._last_signals(x, y)
            f = self.functions.get(x, y)
            o = f.apply(inputs) if f is not None else [0.0] * self.n_of_outputs(x, y)
            out.set(x, y, float(o[0]))
            cur.set(x, y, [float(v) for v in o[1:]])
        for (x, y), c in cur:
            if c is not None:
                self.last_signals.set(x, y, c)
        return out

    def params(self, voxels, tick_times):
        ws = []
        for (x, y), v in voxels:
            if v is None:
                continue
            f = self.functions.get(x, y)
            if not isinstance(f, MultiLayerPerceptron):
                raise NotImplementedError("DistributedSensing: the accelerated path needs a MultiLayerPerceptron "
                                          f"for every voxel (none at {x},{y})")
            if f.get_input_dimension() != self.n_of_inputs(x, y) or f.get_output_dimension() != self.n_of_outputs(x, y):
                raise ValueError(f"Wrong dimension of input or output in the function of voxel ({x},{y})")
            ws.append(np.asarray(f.weights, dtype=np.float64))
        return np.concatenate(ws)


class DistributedSensingNonDirectional(DistributedSensing):
    """controllers/DistributedSensingNonDirectional.java: ``1 + signals`` outputs; every neighbour reads the
    same ``signals`` values (:57-78)."""
    KIND = A.VSR_CTRL_DISTRIBUTED_ND

    @staticmethod
    def n_of_outputs_of(voxel: "Voxel", signals: int) -> int:  # :57-59
        return 1 + signals

    def _last_signals(self, x: int, y: int):  # :61-78
        vals = [0.0] * (4 * self.signals)
        c = 0
        for dx, dy, _ in self.DIRS:
            ax, ay = x + dx, y + dy
            adj = self.last_signals.get(ax, ay) if 0 <= ax < self.last_signals.w and 0 <= ay < self.last_signals.h else None
            if adj is not None:
                vals[c:c + self.signals] = adj[0:self.signals]
            c += self.signals
        return vals


class Robot:
    """core/objects/Robot.java:57-64 (controller + Grid<Voxel>)."""

    def __init__(self, controller: Controller, voxels: Grid):
        self.controller, self.voxels = controller, voxels

    def get_voxels(self) -> Grid:
        return self.voxels

    def get_controller(self) -> Controller:
        return self.controller


# --------------------------------------------------------------------------- RobotUtils
def _params(pattern: str, string: str):
    """util/Utils.java:147-164 (named groups of a full match, else None)."""
    m = re.fullmatch(pattern.replace("(?<", "(?P<"), string)
    return None if m is None else m.groupdict()


class RobotUtils:
    """util/RobotUtils.java: body-shape and sensor-layout mini DSLs."""

    @staticmethod
    def build_robot_transformation(name: str, external_random: Optional["JavaRandom"] = None):
        """RobotUtils.buildRobotTransformation (:65-123): ``identity``, ``breakable-<time|area>-<thr mean>/<thr sd>-
        <restore mean>/<restore sd>-<seed|rnd>`` (every voxel becomes a BreakableVoxel whose actuator freezes) and
        ``broken-<ratio>-<seed|rnd>`` (a share of the voxels is broken from the first tick on, for good).  The draws
        follow the reference's order: per voxel in Grid order nextInt, nextGaussian (threshold), nextGaussian (restore
        time) / nextDouble, nextInt."""
        breakable = (r"breakable-(?<triggerType>time|area)-(?<thresholdMean>\d+(\.\d+)?)/(?<thresholdStDev>\d+(\.\d+)?)-"
                     r"(?<restTimeMean>\d+(\.\d+)?)/(?<restTimeStDev>\d+(\.\d+)?)-(?<seed>\d+|rnd)")
        broken = r"broken-(?<ratio>\d+(\.\d+)?)-(?<seed>\d+|rnd)"
        if _params("identity", name) is not None:
            return lambda robot: robot

        def rnd_of(p):
            if p["seed"] != "rnd":
                return JavaRandom(int(p["seed"]))
            if external_random is None:
                raise ValueError("seed 'rnd' needs an external random generator")
            return external_random

        p = _params(breakable, name)
        if p is not None:
            rnd = rnd_of(p)
            tm, ts = float(p["thresholdMean"]), float(p["thresholdStDev"])
            rm, rs = float(p["restTimeMean"]), float(p["restTimeStDev"])
            trig = p["triggerType"].upper()

            def mk(v):
                if v is None:
                    return None
                seed = rnd.next_int()
                thr = rnd.next_gaussian() * ts + tm
                return BreakableVoxel(v.get_sensors(), seed, {"ACTUATOR": {"FROZEN"}}, {trig: thr}, rnd.next_gaussian() * rs + rm)
            return lambda robot: Robot(robot.get_controller(), Grid.create_from(robot.get_voxels(), mk))
        p = _params(broken, name)
        if p is not None:
            rnd = rnd_of(p)
            ratio = float(p["ratio"])

            def mk(v):
                if v is None:
                    return None
                if rnd.next_double() > ratio:
                    return v
                return BreakableVoxel(v.get_sensors(), rnd.next_int(), {"ACTUATOR": {"FROZEN"}}, {"TIME": 0.0}, math.inf)
            return lambda robot: Robot(robot.get_controller(), Grid.create_from(robot.get_voxels(), mk))
        raise ValueError(f"Unknown transformation name: {name}")

    PREDEFINED_SENSORS = {  # :38-60 (TreeMap => sorted keys); the lidars "l1", "l5" are not on the device
        "a": lambda x, y: SoftNormalization(AreaRatio()),
        "cpg": lambda x, y: Normalization(TimeFunction.sin(-1.0)),
        "l1": lambda x, y: Normalization(Lidar.from_sides(10.0, {RobotUtils.lidar_side(x, y): 1})),
        "l5": lambda x, y: Normalization(Lidar.from_sides(10.0, {RobotUtils.lidar_side(x, y): 5})),
        "m": lambda x, y: Malfunction(),
        "px": lambda x, y: Constant(x),
        "py": lambda x, y: Constant(y),
        "r": lambda x, y: SoftNormalization(Angle()),
        "ax": lambda x, y: SoftNormalization(Trend(Velocity(True, 4.0, Velocity.X), 0.5)),
        "axy": lambda x, y: SoftNormalization(Trend(Velocity(True, 4.0, Velocity.X, Velocity.Y), 0.5)),
        "ay": lambda x, y: SoftNormalization(Trend(Velocity(True, 4.0, Velocity.Y), 0.5)),
        "t": lambda x, y: Average(Touch(), 0.25),
        "vx": lambda x, y: SoftNormalization(Average(Velocity(True, 8.0, Velocity.X), 0.5)),
        "vxy": lambda x, y: SoftNormalization(Average(Velocity(True, 8.0, Velocity.X, Velocity.Y), 0.5)),
        "vy": lambda x, y: SoftNormalization(Average(Velocity(True, 8.0, Velocity.Y), 0.5)),
    }

    @staticmethod
    def lidar_side(x: float, y: float) -> str:  # :244-255
        if y > x and y > (1 - x):
            return "N"
        if y >= x and y <= (1 - x):
            return "W"
        if y < x and y < (1 - x):
            return "S"
        return "E"

    @staticmethod
    def build_shape(name: str) -> Grid:  # :198-242
        p = _params(r"(box|worm)-(?<w>\d+)x(?<h>\d+)", name)
        if p is not None:
            return Grid.create(int(p["w"]), int(p["h"]), True)
        p = _params(r"biped-(?<w>\d+)x(?<h>\d+)", name)
        if p is not None:
            w, h = int(p["w"]), int(p["h"])
            return Grid.create(w, h, lambda x, y: not (y < h // 2 and x >= w // 4 and x < w * 3 // 4))
        p = _params(r"tripod-(?<w>\d+)x(?<h>\d+)", name)
        if p is not None:
            w, h = int(p["w"]), int(p["h"])
            return Grid.create(w, h, lambda x, y: not (y < h // 2 and x != 0 and x != w - 1 and x != w // 2))
        p = _params(r"ball-(?<d>\d+)", name)
        if p is not None:
            d = int(p["d"])
            c = (d - 1) / 2.0

            def inside(x, y):  # Math.round = floor(v + 0.5)
                return math.floor(math.sqrt((x - c) ** 2 + (y - c) ** 2) + 0.5) <= int(math.floor(d / 2.0))

            return Grid.create(d, d, inside)
        p = _params(r"comb-(?<w>\d+)x(?<h>\d+)", name)
        if p is not None:
            w, h = int(p["w"]), int(p["h"])
            return Grid.create(w, h, lambda x, y: (y >= h // 2 or x % 2 == 0))
        p = _params(r"t-(?<w>\d+)x(?<h>\d+)", name)
        if p is not None:
            w, h = int(p["w"]), int(p["h"])
            pad = int(math.floor(math.floor(w / 2.0) / 2.0))
            return Grid.create(w, h, lambda x, y: (y == 0 or (x >= pad and x < h - pad - 1)))
        raise ValueError(f"Unknown body name: {name}")

    @staticmethod
    def sensor(name: str, x: int, y: int, body: Grid) -> Sensor:  # :249-263
        w1, h1 = body.w - 1.0, body.h - 1.0
        nx = x / w1 if w1 != 0 else math.nan
        ny = y / h1 if h1 != 0 else math.nan
        return RobotUtils.PREDEFINED_SENSORS[name](nx, ny)

    @staticmethod
    def build_sensorizing_function(name: str) -> Callable[[Grid], Grid]:  # :124-196
        keys = "|".join(sorted(RobotUtils.PREDEFINED_SENSORS))
        p = _params(rf"uniform-(?<sensors>({keys})(\+({keys}))*)-(?<noiseSigma>\d+(\.\d+)?)", name)
        if p is not None:
            sigma = float(p["noiseSigma"])
            names = p["sensors"].split("+")
            noisy = (lambda s: s) if sigma == 0 else (lambda s: Noisy(s, sigma, 0))  # :172-175
            return lambda body: Grid.create(
                body.w, body.h,
                lambda x, y: None if not body.get(x, y) else Voxel([noisy(RobotUtils.sensor(n, x, y, body)) for n in names]),
            )
        if _params("empty", name) is not None:
            return lambda body: Grid.create(body.w, body.h, lambda x, y: None if not body.get(x, y) else Voxel([]))
        p = _params(r"spinedTouch(?<sighted>Sighted)?-(?<cpg>[tf])-(?<malfunction>[tf])-(?<noiseSigma>\d+(\.\d+)?)", name)
        if p is not None:  # :134-165
            sighted = p["sighted"] is not None
            sg = float(p["noiseSigma"])

            def spined(body, x, y, p=p, sg=sg):
                pick = [("a", True), ("m", p["malfunction"] == "t"), ("t", y == 0), ("vxy", y == body.h - 1),
                        ("cpg", x == body.w - 1 and y == body.h - 1 and p["cpg"] == "t"),
                        ("l5", sighted and x == body.w - 1)]
                ss = [RobotUtils.sensor(n, x, y, body) for n, cond in pick if cond]
                return Voxel(ss if sg == 0 else [Noisy(s, sg, 0) for s in ss])

            return lambda body: Grid.create(body.w, body.h, lambda x, y: None if not body.get(x, y) else spined(body, x, y))
        p = _params(r"uniformAll-(?<noiseSigma>\d+(\.\d+)?)", name)
        if p is not None:  # :177-189: every predefined sensor, TreeMap (sorted) order
            return RobotUtils.build_sensorizing_function(
                "uniform-" + "+".join(sorted(RobotUtils.PREDEFINED_SENSORS)) + "-" + p["noiseSigma"])
        raise ValueError(f"Unknown sensorizing function name: {name}")


# --------------------------------------------------------------------------- Settings / Outcome
class Settings:
    """org.dyn4j.dynamics.Settings as a plain value holder (Locomotion.java:50)."""

    def __init__(self):
        self._s = A.default_settings()

    def get_step_frequency(self) -> float:
        return self._s.step_frequency

    def __getattr__(self, k):
        return getattr(self.__dict__["_s"], k)

    def set(self, **kw) -> "Settings":
        for k, v in kw.items():
            setattr(self._s, k, v)
        return self

    def raw(self) -> A.vsr_settings:
        return self._s


class Snapshot:
    """core/snapshots/Snapshot.java: content + the class that produced it + children."""

    def __init__(self, content, snapshottable_class: str, children=None):
        self.content, self.snapshottable_class = content, snapshottable_class
        self.children = list(children) if children else []

    def get_content(self):
        return self.content

    def get_children(self):
        return self.children

    @staticmethod
    def world(snapshots) -> "Snapshot":  # Snapshot.world(List<Snapshot>)
        return Snapshot(None, "World", snapshots)


class RobotShape:
    """core/snapshots/RobotShape.java: Grid<VoxelPoly> + bounding box."""

    def __init__(self, polies: "Grid", bounding_box):
        self.polies, self.box = polies, bounding_box

    def get_polies(self) -> "Grid":
        return self.polies

    def bounding_box(self):
        return self.box


class Outcome:
    """tasks/locomotion/Outcome.java.

    Built either from the ``vsr_result`` record (what the first and the last Observation give: distance,
    time, velocity, energies -- the fitness path, nothing per tick leaves the device) or from the full
    ``observations`` map (``Locomotion.apply(..., observations=True)`` or with a listener), which adds
    the spectra, postures, footprints and gaits."""

    def __init__(self, rec=None, observations=None):
        if isinstance(rec, dict) and observations is None:  # new Outcome(Map<Double, Observation>)
            rec, observations = None, rec
        self.rec = rec
        self.observations = dict(sorted(observations.items())) if observations is not None else None

    def get_status(self) -> int:
        """``vsr_result.status`` (include/vsr_b200.h, VSR_STATUS_*): 0 = a valid episode.  Not in the reference
        (dyn4j has no contact capacity and lets NaN through); ``Locomotion.apply`` raises on a non-zero
        status unless told otherwise."""
        return 0 if self.rec is None else int(self.rec["status"])

    # ---- first / last observation
    def _ends(self):
        ks = list(self.observations)
        return self.observations[ks[0]], self.observations[ks[-1]], ks[0], ks[-1]

    def _polies(self, o):
        return [v for v in o.voxel_polies.values() if v is not None]

    def get_distance(self) -> float:  # :152-166
        if self.observations is None:
            return float(self.rec["last_center_x"] - self.rec["first_center_x"])
        f, l, _, _ = self._ends()
        from .behavior import BehaviorUtils
        return BehaviorUtils.center(self._polies(l))[0] - BehaviorUtils.center(self._polies(f))[0]

    def get_time(self) -> float:  # :194-196
        if self.observations is None:
            return float(self.rec["t_last"] - self.rec["t_first"])
        _, _, t0, t1 = self._ends()
        return t1 - t0

    def get_velocity(self) -> float:  # :198-200
        return self.get_distance() / self.get_time()

    def get_area_ratio_energy(self) -> float:  # :42-58
        if self.observations is None:
            return float(self.rec["area_ratio_energy_last"] - self.rec["area_ratio_energy_first"])
        f, l, _, _ = self._ends()
        return sum(v.get_area_ratio_energy() for v in self._polies(l)) - sum(v.get_area_ratio_energy() for v in self._polies(f))

    def get_control_energy(self) -> float:  # :126-142
        if self.observations is None:
            return float(self.rec["control_energy_last"] - self.rec["control_energy_first"])
        f, l, _, _ = self._ends()
        return sum(v.get_control_energy() for v in self._polies(l)) - sum(v.get_control_energy() for v in self._polies(f))

    def get_control_power(self) -> float:
        return self.get_control_energy() / self.get_time()

    def get_area_ratio_power(self) -> float:
        return self.get_area_ratio_energy() / self.get_time()

    def get_corrected_efficiency(self) -> float:  # :148-150
        return self.get_distance() / (1.0 + self.get_control_power() * self.get_time())

    # ---- the per-tick part
    def _need(self):
        if self.observations is None:
            raise NotImplementedError("this Outcome holds the first/last observation only: "
                                      "run Locomotion.apply(..., observations=True)")
        return self.observations

    def get_observations(self):  # :190-192
        return self._need()

    def get_computation_time(self) -> float:  # :121-124
        f, l, _, _ = (self._need(), *self._ends())[1:]
        return l.computation_time - f.computation_time

    def sub_outcome(self, start_t: float, end_t: float) -> "Outcome":  # :202-204: subMap(startT, endT)
        return Outcome(observations={t: o for t, o in self._need().items() if start_t <= t < end_t})

    def get_average_posture(self, n: int):  # :64-69
        from .behavior import BehaviorUtils
        return BehaviorUtils.compute_average_posture(
            [BehaviorUtils.compute_posture(self._polies(o), n) for o in self._need().values()])

    def _center_signal(self, fn):
        from .behavior import BehaviorUtils
        return {t: fn(BehaviorUtils.get_central_element(o.voxel_polies)) for t, o in self._need().items()}

    def _spectrum(self, fn, min_f, max_f, n_bins):
        from .behavior import BehaviorUtils
        return BehaviorUtils.compute_quantized_spectrum(self._center_signal(fn), min_f, max_f, n_bins)

    def get_center_angle_spectrum(self, min_f, max_f, n_bins):  # :71-79
        return self._spectrum(lambda v: v.get_angle(), min_f, max_f, n_bins)

    def get_center_x_position_spectrum(self, min_f, max_f, n_bins):  # :81-89
        return self._spectrum(lambda v: v.center()[0], min_f, max_f, n_bins)

    def get_center_x_velocity_spectrum(self, min_f, max_f, n_bins):  # :91-99
        return self._spectrum(lambda v: v.get_linear_velocity()[0], min_f, max_f, n_bins)

    def get_center_y_position_spectrum(self, min_f, max_f, n_bins):  # :101-109
        return self._spectrum(lambda v: v.center()[1], min_f, max_f, n_bins)

    def get_center_y_velocity_spectrum(self, min_f, max_f, n_bins):  # :111-119
        return self._spectrum(lambda v: v.get_linear_velocity()[1], min_f, max_f, n_bins)

    def get_footprints_spectra(self, n: int, min_f, max_f, n_bins):  # :168-188
        from .behavior import BehaviorUtils
        fps = {t: BehaviorUtils.compute_footprint(self._polies(o), n) for t, o in self._need().items()}
        return [BehaviorUtils.compute_quantized_spectrum({t: (1.0 if f.mask[i] else 0.0) for t, f in fps.items()},
                                                         min_f, max_f, n_bins) for i in range(n)]

    def get_main_gait(self, interval: float, longest_interval: float, n: int):
        """BehaviorUtils.computeMainGait over the observations' voxel polies."""
        from .behavior import BehaviorUtils
        return BehaviorUtils.compute_main_gait(interval, longest_interval,
                                               {t: self._polies(o) for t, o in self._need().items()}, n)

    def __repr__(self):
        return f"Outcome{{distance={self.get_distance():.6f}, time={self.get_time():.3f}, velocity={self.get_velocity():.6f}}}"


RESULT_DTYPE = np.dtype([
    ("first_center_x", "<f8"), ("first_center_y", "<f8"), ("last_center_x", "<f8"), ("last_center_y", "<f8"),
    ("t_first", "<f8"), ("t_last", "<f8"),
    ("area_ratio_energy_first", "<f8"), ("area_ratio_energy_last", "<f8"),
    ("control_energy_first", "<f8"), ("control_energy_last", "<f8"),
    ("n_ticks", "<i4"), ("status", "<i4"),
])
assert RESULT_DTYPE.itemsize == C.sizeof(A.vsr_result)


def tick_times(final_t: float, dt: float, t_start: float = 0.0) -> np.ndarray:
    """``t = 0; while (t < finalT) t = t + dT`` in double (Locomotion.java:195-203,
    AbstractTask.java:48).  1200 ticks for finalT = 20, dt = 1/60.  ``t_start``: the clock of a later
    stage of a longer task (DevoLocomotion) keeps running."""
    out, t = [], float(t_start)
    while t < final_t:
        t = t + dt
        out.append(t)
    return np.asarray(out, dtype=np.float64)


# --------------------------------------------------------------------------- batch descriptor
class BatchDesc:
    """Owns the flat arrays behind a ``vsr_batch_desc`` (keeps them alive)."""

    def __init__(self):
        self.desc = A.vsr_batch_desc()
        self._keep: list = []
        self.n_voxels: List[int] = []       # per shape
        self.n_readings: List[int] = []     # per shape
        self.robot_shape: np.ndarray = np.zeros(0, np.int32)
        self.n_ticks = 0

    def ptr(self):
        return C.byref(self.desc)

    def voxel_steps(self) -> int:
        nv = np.asarray(self.n_voxels, dtype=np.int64)
        return int(nv[self.robot_shape].sum()) * self.n_ticks

    def keep(self, arr):
        self._keep.append(arr)
        return arr


def _np_ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def compile_batch(
    robots: Sequence[Robot],
    final_t: float,
    ground_profile: Sequence[Sequence[float]],
    initial_placement: Optional[float],
    settings: Settings,
    precision: int = A.VSR_PRECISION_F32,
    trajectory_robots: int = 0,
    t_start: float = 0.0,
    placements: Optional[Sequence[float]] = None,
) -> BatchDesc:
    """Flatten a list of Robots + one Locomotion task into a ``vsr_batch_desc``."""
    if len(robots) == 0:
        raise ValueError("empty batch")
    kinds = {r.controller.KIND for r in robots}
    if len(kinds) != 1:
        raise ValueError("one controller kind per batch")
    kind = kinds.pop()
    b = BatchDesc()
    d = b.desc
    d.abi_version = A.VSR_ABI_VERSION
    d.precision = precision
    d.n_robots = len(robots)
    d.controller_kind = kind
    d.settings = settings.raw()
    d.final_t = final_t
    xs = b.keep(np.ascontiguousarray(ground_profile[0], dtype=np.float64))
    ys = b.keep(np.ascontiguousarray(ground_profile[1], dtype=np.float64))
    if xs.size != ys.size:
        raise ValueError("xs[] and ys[] must have the same length")  # Ground.java:49-51
    d.terrain_xs, d.terrain_ys, d.terrain_n = _np_ptr(xs, C.c_double), _np_ptr(ys, C.c_double), xs.size
    d.initial_placement = float(xs[1] + 1.0) if initial_placement is None else float(initial_placement)
    d.trajectory_robots = trajectory_robots
    d.t_start = float(t_start)
    if placements is not None:
        if len(placements) != len(robots):
            raise ValueError("one placement per robot")
        pl = b.keep(np.ascontiguousarray(placements, dtype=np.float64))
        d.robot_placement = _np_ptr(pl, C.c_double)
    ticks = tick_times(final_t, d.settings.step_frequency, t_start)
    b.n_ticks = len(ticks)

    # shape classes: (mask, per-voxel sensor encodings, hidden layers)
    shape_keys: dict = {}
    shape_brk: dict = {}
    shapes_c: list = []
    robot_shape = np.zeros(len(robots), dtype=np.int32)
    material = None
    activation = None
    dist_signals = None
    plist = []
    body_cache: dict = {}  # robots of a population usually share one body Grid: describe it once
    for i, r in enumerate(robots):
        vox = r.voxels
        cached = body_cache.get(id(vox))
        if cached is None:
            mask = bytes(1 if v is not None else 0 for v in vox.values())
            sens = tuple(
                tuple((s.base, s.axes, s.rotated, s.aggregator, s.soft_norm, s.max_velocity_norm, s.interval, s.noise_sigma,
                       s.noise_seed, tuple(s.lidar_dirs))
                      for s in (ss.encode() for ss in v.get_sensors()))
                for v in vox.values() if v is not None
            )
            mats = {bytes(v.material) for v in vox.values() if v is not None}
            brk = tuple(v.encode_breakable() if isinstance(v, BreakableVoxel) else None for v in vox.values() if v is not None)
            if all(q is None for q in brk):
                brk = None
            cached = body_cache[id(vox)] = (vox, mask, sens, mats, brk)
        _, mask, sens, mats, brk = cached
        hidden: Tuple[int, ...] = ()
        if kind == A.VSR_CTRL_MLP:
            f = r.controller.function
            hidden = tuple(f.neurons[1:-1])
            act = A.VSR_ACT[f.activation]
            if activation is None:
                activation = act
            elif activation != act:
                raise ValueError("one MLP activation per batch")
        elif kind in (A.VSR_CTRL_DISTRIBUTED, A.VSR_CTRL_DISTRIBUTED_ND):
            fs = [r.controller.functions.get(x, y) for (x, y), v in vox if v is not None]
            if any(not isinstance(f, MultiLayerPerceptron) for f in fs):
                raise NotImplementedError("DistributedSensing: the accelerated path needs a MultiLayerPerceptron for every voxel")
            hs = {tuple(f.neurons[1:-1]) for f in fs}
            acts = {A.VSR_ACT[f.activation] for f in fs}
            if len(hs) != 1 or len(acts) != 1:
                raise NotImplementedError("DistributedSensing: one activation and one set of hidden layer sizes per robot")
            hidden, act = hs.pop(), acts.pop()
            if activation is None:
                activation = act
            elif activation != act:
                raise ValueError("one MLP activation per batch")
            if dist_signals is None:
                dist_signals = r.controller.signals
            elif dist_signals != r.controller.signals:
                raise ValueError("one DistributedSensing signal count per batch")
        if len(mats) != 1:
            # (the rows between the voxels of a robot -- welds, mass-mass contacts -- are solved for equal masses)
            raise NotImplementedError("one voxel material per robot on the accelerated path")
        (mat_r,) = mats
        if material is None:
            material = mat_r                      # the batch's default material: the first robot's
        key = (vox.w, vox.h, mask, sens, hidden, repr(brk), mat_r)
        if key not in shape_keys:
            shape_brk[key] = brk
            shape_keys[key] = len(shapes_c)
            shapes_c.append(key)
        robot_shape[i] = shape_keys[key]
        plist.append(np.ascontiguousarray(r.controller.params(vox, ticks), dtype=np.float64))
    if material is None:
        raise ValueError("robot without voxels")
    d.material = A.vsr_material.from_buffer_copy(material)
    d.mlp_activation = activation if activation is not None else 0
    d.distributed_signals = dist_signals if dist_signals is not None else 0

    arr = (A.vsr_shape * len(shapes_c))()
    for k, key_k in enumerate(shapes_c):
        w, h, mask, sens, hidden = key_k[:5]
        m = b.keep(np.frombuffer(mask, dtype=np.uint8).copy())
        if key_k[6] != material:                  # a shape class with its own material (vsr_shape.material)
            mk = b.keep(A.vsr_material.from_buffer_copy(key_k[6]))
            arr[k].material = C.pointer(mk)
        brk = shape_brk.get(key_k)
        if brk is not None:
            barr = (A.vsr_breakable * len(brk))()
            for j, q in enumerate(brk):
                if q is None:
                    continue
                barr[j].enabled = 1
                for z in range(3):
                    barr[j].malfunctions[z] = q[0][z]
                    barr[j].thresholds[z] = q[1][z]
                barr[j].restore_time, barr[j].random_seed = q[2], q[3]
            b.keep(barr)
            arr[k].breakable = C.cast(barr, C.POINTER(A.vsr_breakable))
        offs = np.zeros(len(sens) + 1, dtype=np.int32)
        flat = []
        for j, vs in enumerate(sens):
            flat.extend(vs)
            offs[j + 1] = len(flat)
        b.keep(offs)
        sarr = (A.vsr_sensor * max(1, len(flat)))()
        nread = 0
        for j, t in enumerate(flat):
            s = sarr[j]
            s.base, s.axes, s.rotated, s.aggregator, s.soft_norm, s.max_velocity_norm, s.interval, s.noise_sigma, s.noise_seed = t[:9]
            for z, dv in enumerate(t[9]):
                s.lidar_dirs[z] = dv
            nread += (bin(s.axes & 3).count("1") if s.base == A.VSR_SENSOR_VELOCITY else s.axes if s.base == A.VSR_SENSOR_LIDAR else 1)
        b.keep(sarr)
        hid = b.keep(np.asarray(hidden if hidden else [0], dtype=np.int32))
        arr[k].w, arr[k].h = w, h
        arr[k].mask = _np_ptr(m, C.c_uint8)
        arr[k].sensor_offsets = _np_ptr(offs, C.c_int32)
        arr[k].sensors = C.cast(sarr, C.POINTER(A.vsr_sensor))
        arr[k].n_hidden = len(hidden)
        arr[k].hidden = _np_ptr(hid, C.c_int32)
        b.n_voxels.append(len(sens))
        b.n_readings.append(nread)
    b.keep(arr)
    d.shapes, d.n_shapes = C.cast(arr, C.POINTER(A.vsr_shape)), len(shapes_c)
    b.robot_shape = b.keep(robot_shape)
    d.robot_shape = _np_ptr(robot_shape, C.c_int32)
    offs = np.zeros(len(robots) + 1, dtype=np.uint64)
    np.cumsum([p.size for p in plist], out=offs[1:])
    params = b.keep(np.concatenate(plist) if plist else np.zeros(1))
    b.keep(offs)
    d.params = _np_ptr(params, C.c_double)
    d.param_offsets = C.cast(offs.ctypes.data, C.POINTER(C.c_size_t))
    return b


class VsrError(RuntimeError):
    pass


def _check(lib, rc: int) -> None:
    if rc == 0:
        return
    msg = lib.vsr_last_error().decode()
    if rc in (A.VSR_ERR_INVALID_ARGUMENT,):
        raise ValueError(msg)
    if rc == A.VSR_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise VsrError(f"vsr error {rc}: {msg}")


class DeviceBatch:
    """RAII wrapper of a ``vsr_batch*``."""

    def __init__(self, bd: BatchDesc):
        self.lib = A.load_library()
        self.bd = bd
        self.h = C.c_void_p()
        _check(self.lib, self.lib.vsr_batch_create(bd.ptr(), C.byref(self.h)))

    def upload(self, device: int = 0, stream: int = 0):
        _check(self.lib, self.lib.vsr_batch_upload(self.h, device, C.c_void_p(stream)))
        return self

    def run(self, stream: int = 0):
        _check(self.lib, self.lib.vsr_batch_run(self.h, C.c_void_p(stream)))
        return self

    def results(self) -> np.ndarray:
        n = int(self.bd.desc.n_robots)
        out = np.zeros(n, dtype=RESULT_DTYPE)
        _check(self.lib, self.lib.vsr_batch_results(self.h, C.cast(out.ctypes.data, C.POINTER(A.vsr_result)), n))
        return out

    def results_device(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        _check(self.lib, self.lib.vsr_batch_results_device(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def trajectory(self, robot: int) -> np.ndarray:
        nb = self.lib.vsr_batch_n_bodies(self.h, robot)
        nt = self.lib.vsr_batch_n_ticks(self.h)
        out = np.zeros((nt, nb, 6), dtype=np.float64)
        got = self.lib.vsr_batch_trajectory(self.h, robot, _np_ptr(out, C.c_double), out.size)
        if got < 0:
            _check(self.lib, int(got))
        return out

    def final_bbox(self) -> np.ndarray:
        """[robot][4] = Robot.boundingBox at the last tick: min x, min y, max x, max y (vsr_batch_final_bbox)."""
        n = int(self.bd.desc.n_robots)
        out = np.zeros((n, 4), dtype=np.float64)
        _check(self.lib, self.lib.vsr_batch_final_bbox(self.h, _np_ptr(out, C.c_double), n))
        return out

    def voxel_trace(self, robot: int) -> np.ndarray:
        """[tick][voxel][2] = touching the ground (0/1), applied force (vsr_batch_voxel_trace)."""
        nv = self.lib.vsr_batch_n_bodies(self.h, robot) // 4
        nt = self.lib.vsr_batch_n_ticks(self.h)
        out = np.zeros((nt, nv, 2), dtype=np.float64)
        got = self.lib.vsr_batch_voxel_trace(self.h, robot, _np_ptr(out, C.c_double), out.size)
        if got < 0:
            _check(self.lib, int(got))
        return out

    def sensor_readings(self, robot: int) -> np.ndarray:
        """[tick][n_readings]: Voxel.getSensorReadings of the robot's voxels, concatenated in Grid.values() order
        (vsr_batch_sensor_readings)."""
        nt = self.lib.vsr_batch_n_ticks(self.h)
        cap = nt * 24 * (self.lib.vsr_batch_n_bodies(self.h, robot) // 4)
        out = np.zeros(cap, dtype=np.float64)
        got = self.lib.vsr_batch_sensor_readings(self.h, robot, _np_ptr(out, C.c_double), out.size)
        if got < 0:
            _check(self.lib, int(got))
        return out[:got].reshape(nt, got // nt if nt else 0)

    def voxel_steps(self) -> int:
        return int(self.lib.vsr_batch_voxel_steps(self.h))

    def last_launches(self) -> int:
        return int(self.lib.vsr_batch_last_launches(self.h))

    def close(self):
        if self.h:
            self.lib.vsr_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- Locomotion
class Locomotion:
    """tasks/locomotion/Locomotion.java.  ``apply`` takes one Robot or a list of them
    (the batch entry the reference lacks: ``List<Outcome> apply(List<Robot>)``)."""

    INITIAL_PLACEMENT_X_GAP = 1.0
    INITIAL_PLACEMENT_Y_GAP = 1.0
    TERRAIN_BORDER_HEIGHT = 100.0
    TERRAIN_LENGTH = 2000
    TERRAIN_BORDER_WIDTH = 10.0

    def __init__(self, final_t: float, ground_profile, settings: Settings, initial_placement: Optional[float] = None):
        self.final_t = final_t
        self.ground_profile = [list(map(float, ground_profile[0])), list(map(float, ground_profile[1]))]
        xs, ys = self.ground_profile
        if len(xs) != len(ys):
            raise ValueError("xs[] and ys[] must have the same length")  # Ground.java:49-51
        if len(xs) < 2:
            raise ValueError("There must be at least 2 points")  # Ground.java:52-54
        if sorted(xs) != xs:
            raise ValueError("x coordinates must be sorted")  # Ground.java:55-58
        # :50-52: groundProfile[0][1] + INITIAL_PLACEMENT_X_GAP
        self.initial_placement = xs[1] + self.INITIAL_PLACEMENT_X_GAP if initial_placement is None else initial_placement
        self.settings = settings

    @staticmethod
    def create_terrain(name: str):  # :61-146
        L, BW, BH = Locomotion.TERRAIN_LENGTH, Locomotion.TERRAIN_BORDER_WIDTH, Locomotion.TERRAIN_BORDER_HEIGHT
        num = r"[0-9]+(\.[0-9]+)?"
        if _params("flat", name) is not None:
            return [[0.0, BW, L - BW, float(L)], [BH, 5.0, 5.0, BH]]
        p = _params(r"flatWithStart-(?<seed>[0-9]+)", name)
        if p is not None:
            rnd = JavaRandom(int(p["seed"]))
            for _ in range(rnd.next_int(10) + 10):
                rnd.next_double()
            angle = math.pi / 18.0 * (rnd.next_double() * 2.0 - 1.0)
            start = Voxel.SIDE_LENGTH * 8.0
            return [[0.0, BW, BW + start, L - BW, float(L)], [BH, 5 + start * math.sin(angle), 5.0, 5.0, BH]]
        p = _params(rf"hilly-(?<h>{num})-(?<w>{num})-(?<seed>[0-9]+)", name)
        if p is not None:
            h, w = float(p["h"]), float(p["w"])
            rnd = JavaRandom(int(p["seed"]))
            xs, ys = [0.0, BW], [BH, 0.0]
            while xs[-1] < L - BW:
                xs.append(xs[-1] + max(1.0, (rnd.next_gaussian() * 0.25 + 1) * w))
                ys.append(ys[-1] + rnd.next_gaussian() * h)
            xs.append(xs[-1] + BW)
            ys.append(BH)
            return [xs, ys]
        p = _params(rf"steppy-(?<h>{num})-(?<w>{num})-(?<seed>[0-9]+)", name)
        if p is not None:
            h, w = float(p["h"]), float(p["w"])
            rnd = JavaRandom(int(p["seed"]))
            xs, ys = [0.0, BW], [BH, 0.0]
            while xs[-1] < L - BW:
                xs.append(xs[-1] + max(1.0, (rnd.next_gaussian() * 0.25 + 1) * w))
                xs.append(xs[-1] + 0.5)
                ys.append(ys[-1])
                ys.append(ys[-1] + rnd.next_gaussian() * h)
            xs.append(xs[-1] + BW)
            ys.append(BH)
            return [xs, ys]
        p = _params(rf"downhill-(?<angle>{num})", name)
        if p is not None:
            dy = (L - 2 * BW) * math.sin(float(p["angle"]) / 180 * math.pi)
            return [[0.0, BW, L - BW, float(L)], [BH + dy, 5 + dy, 5.0, BH]]
        p = _params(rf"uphill-(?<angle>{num})", name)
        if p is not None:
            dy = (L - 2 * BW) * math.sin(float(p["angle"]) / 180 * math.pi)
            return [[0.0, BW, L - BW, float(L)], [BH, 5.0, 5 + dy, BH + dy]]
        raise ValueError(f"Unknown terrain name: {name}")

    def compile(self, robots: Sequence[Robot], precision: int = A.VSR_PRECISION_F32,
                trajectory_robots: int = 0, t_start: float = 0.0, placements=None) -> BatchDesc:
        return compile_batch(robots, self.final_t, self.ground_profile, self.initial_placement, self.settings,
                             precision, trajectory_robots, t_start, placements)

    def apply(self, robots, listener=None, device: int = 0, precision: int = A.VSR_PRECISION_F32,
              observations: bool = False, check_status: bool = True):
        """``Outcome apply(Robot[, SnapshotListener])`` (Locomotion.java:168-207) and its batched form
        ``List<Outcome> apply(List<Robot>)``: one kernel launch for all robots.

        Without a listener and with ``observations=False`` only the ``vsr_result`` records come back
        (the fitness path).  ``observations=True`` also records every tick on the device and returns
        Outcomes holding the reference's ``observations`` map; a ``listener`` (an object with
        ``listen(t, snapshot)`` or a callable) is fed ``Snapshot.world([ground, robot])`` per tick and
        robot, after the episode (the device runs all ticks in one launch)."""
        import time as _time
        single = isinstance(robots, Robot)
        rs = [robots] if single else list(robots)
        full = observations or listener is not None
        bd = self.compile(rs, precision, trajectory_robots=len(rs) if full else 0)
        t0 = _time.time()
        with DeviceBatch(bd) as db:
            db.upload(device).run()
            res = db.results()
            dumps = [(db.trajectory(i), db.voxel_trace(i)) for i in range(len(rs))] if full else None
        wall = _time.time() - t0
        bad = np.flatnonzero(res["status"] != 0)
        if check_status and bad.size:
            # truncated physics (contact capacity) or a non-finite state must not come back as a silent Outcome
            raise RuntimeError(f"{bad.size} robot(s) ended with a non-zero status (first: robot {int(bad[0])}, status "
                               f"{int(res['status'][bad[0]])}; bits: 1 non-finite, 2/4/8 contact capacity, see "
                               "include/vsr_b200.h); pass check_status=False to get the flagged Outcomes")
        if not full:
            outs = [Outcome(res[i]) for i in range(len(rs))]
            return outs[0] if single else outs
        from .behavior import build_observations
        tt = tick_times(self.final_t, self.settings.get_step_frequency())
        outs = []
        for i, r in enumerate(rs):
            mat = next(v for v in r.get_voxels().values() if v is not None).material
            obs = build_observations(r.get_voxels(), dumps[i][0], dumps[i][1], tt, (self.ground_profile[0], self.ground_profile[1]),
                                     mat.side_length, mat.mass_side_length_ratio * mat.side_length, wall)
            outs.append(Outcome(res[i], obs))
            if listener is not None:
                ground = Snapshot((list(self.ground_profile[0]), list(self.ground_profile[1])), "Ground")
                for t, o in obs.items():
                    polies = [v for v in o.voxel_polies.values() if v is not None]
                    from .behavior import BoundingBox
                    shape = RobotShape(o.voxel_polies, BoundingBox.largest(p.bounding_box() for p in polies))
                    snap = Snapshot.world([ground, Snapshot(shape, "Robot", [Snapshot(p, "Voxel") for p in polies])])
                    (listener.listen if hasattr(listener, "listen") else listener)(t, snap)
        return outs[0] if single else outs
