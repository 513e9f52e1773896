"""ctypes mirror of include/vsr_b200.h (the C-ABI drop-in boundary).

Only struct layouts, constants and the loader for the product library
``libvsr_b200.so`` live here.  There is no CPU fallback: if the CUDA library is
missing, :func:`load_library` raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvsr_b200.so")

VSR_ABI_VERSION = 2

VSR_OK = 0
VSR_ERR_INVALID_ARGUMENT = -1
VSR_ERR_UNSUPPORTED = -2
VSR_ERR_CUDA = -3
VSR_ERR_STATE = -4
VSR_ERR_NO_DEVICE = -5

VSR_SPRING_MASS_REDUCED = 0
VSR_SPRING_MASS_EFFECTIVE = 1

VSR_SCAFFOLD_SIDE_EXTERNAL = 1
VSR_SCAFFOLD_SIDE_INTERNAL = 2
VSR_SCAFFOLD_SIDE_CROSS = 4
VSR_SCAFFOLD_CENTRAL_CROSS = 8
VSR_SCAFFOLD_ALL = 15

VSR_SENSOR_TOUCH = 0
VSR_SENSOR_AREA_RATIO = 1
VSR_SENSOR_VELOCITY = 2
VSR_AGG_NONE = 0
VSR_AGG_AVERAGE = 1
VSR_AGG_TREND = 2
VSR_AGG_DYNAMIC_NORM = 3
VSR_AXIS_X = 1
VSR_AXIS_Y = 2

VSR_CTRL_PHASE_SIN = 0
VSR_CTRL_TABULATED = 1
VSR_SENSOR_ANGLE, VSR_SENSOR_CONSTANT, VSR_SENSOR_APPLIED_FORCE, VSR_SENSOR_MALFUNCTION, VSR_SENSOR_TIME_SIN = 3, 4, 5, 6, 7
VSR_SENSOR_CRUMPLING, VSR_SENSOR_CONTROL_POWER, VSR_SENSOR_LIDAR = 8, 9, 10
VSR_LIDAR_MAX_RAYS = 8
VSR_NORM_NONE, VSR_NORM_SOFT, VSR_NORM_LINEAR = 0, 1, 2
VSR_CTRL_MLP = 2
VSR_CTRL_DISTRIBUTED = 3
VSR_CTRL_DISTRIBUTED_ND = 4

VSR_ACT = {"RELU": 0, "SIGMOID": 1, "SIN": 2, "TANH": 3, "SIGN": 4, "IDENTITY": 5}

VSR_PRECISION_F32 = 0
VSR_PRECISION_F64 = 1


class vsr_settings(C.Structure):
    _fields_ = [
        ("step_frequency", C.c_double),
        ("velocity_iterations", C.c_int32),
        ("position_iterations", C.c_int32),
        ("linear_tolerance", C.c_double),
        ("angular_tolerance", C.c_double),
        ("max_linear_correction", C.c_double),
        ("baumgarte", C.c_double),
        ("restitution_velocity", C.c_double),
        ("max_translation", C.c_double),
        ("max_rotation", C.c_double),
        ("gravity_x", C.c_double),
        ("gravity_y", C.c_double),
        ("ground_friction", C.c_double),
        ("ground_restitution", C.c_double),
        ("spring_mass_mode", C.c_int32),
        ("self_collision", C.c_int32),
        ("continuous_detection", C.c_int32),
        ("reserved", C.c_int32),
    ]


class vsr_material(C.Structure):
    _fields_ = [
        ("side_length", C.c_double),
        ("mass_side_length_ratio", C.c_double),
        ("spring_f", C.c_double),
        ("spring_d", C.c_double),
        ("mass_linear_damping", C.c_double),
        ("mass_angular_damping", C.c_double),
        ("friction", C.c_double),
        ("restitution", C.c_double),
        ("mass", C.c_double),
        ("area_ratio_passive_min", C.c_double),
        ("area_ratio_passive_max", C.c_double),
        ("area_ratio_active_min", C.c_double),
        ("area_ratio_active_max", C.c_double),
        ("spring_scaffoldings", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class vsr_sensor(C.Structure):
    _fields_ = [
        ("base", C.c_int32),
        ("axes", C.c_int32),
        ("rotated", C.c_int32),
        ("aggregator", C.c_int32),
        ("soft_norm", C.c_int32),
        ("reserved", C.c_int32),
        ("max_velocity_norm", C.c_double),
        ("interval", C.c_double),
        ("noise_sigma", C.c_double),
        ("noise_seed", C.c_int64),
        ("lidar_dirs", C.c_double * VSR_LIDAR_MAX_RAYS),
    ]


VSR_MALFUNCTION_NONE, VSR_MALFUNCTION_ZERO, VSR_MALFUNCTION_FROZEN, VSR_MALFUNCTION_RANDOM = 0, 1, 2, 3
VSR_COMPONENT_ACTUATOR, VSR_COMPONENT_SENSORS, VSR_COMPONENT_STRUCTURE = 0, 1, 2
VSR_TRIGGER_CONTROL, VSR_TRIGGER_AREA, VSR_TRIGGER_TIME = 0, 1, 2


class vsr_breakable(C.Structure):
    _fields_ = [
        ("enabled", C.c_int32),
        ("malfunctions", C.c_uint32 * 3),
        ("thresholds", C.c_double * 3),
        ("restore_time", C.c_double),
        ("random_seed", C.c_int64),
    ]


class vsr_shape(C.Structure):
    _fields_ = [
        ("w", C.c_int32),
        ("h", C.c_int32),
        ("mask", C.POINTER(C.c_uint8)),
        ("sensor_offsets", C.POINTER(C.c_int32)),
        ("sensors", C.POINTER(vsr_sensor)),
        ("n_hidden", C.c_int32),
        ("reserved", C.c_int32),
        ("hidden", C.POINTER(C.c_int32)),
        ("breakable", C.POINTER(vsr_breakable)),
        ("material", C.POINTER(vsr_material)),
    ]


class vsr_batch_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("precision", C.c_uint32),
        ("n_robots", C.c_size_t),
        ("n_shapes", C.c_int32),
        ("controller_kind", C.c_int32),
        ("shapes", C.POINTER(vsr_shape)),
        ("robot_shape", C.POINTER(C.c_int32)),
        ("material", vsr_material),
        ("mlp_activation", C.c_int32),
        ("distributed_signals", C.c_int32),
        ("params", C.POINTER(C.c_double)),
        ("param_offsets", C.POINTER(C.c_size_t)),
        ("final_t", C.c_double),
        ("terrain_xs", C.POINTER(C.c_double)),
        ("terrain_ys", C.POINTER(C.c_double)),
        ("terrain_n", C.c_int32),
        ("reserved1", C.c_int32),
        ("initial_placement", C.c_double),
        ("settings", vsr_settings),
        ("trajectory_robots", C.c_size_t),
        ("t_start", C.c_double),
        ("robot_placement", C.POINTER(C.c_double)),
    ]


class vsr_result(C.Structure):
    _fields_ = [
        ("first_center_x", C.c_double),
        ("first_center_y", C.c_double),
        ("last_center_x", C.c_double),
        ("last_center_y", C.c_double),
        ("t_first", C.c_double),
        ("t_last", C.c_double),
        ("area_ratio_energy_first", C.c_double),
        ("area_ratio_energy_last", C.c_double),
        ("control_energy_first", C.c_double),
        ("control_energy_last", C.c_double),
        ("n_ticks", C.c_int32),
        ("status", C.c_int32),
    ]


RESULT_DOUBLES = C.sizeof(vsr_result) // 8  # 11 (10 doubles + 2 int32)


def default_settings() -> vsr_settings:
    """dyn4j 4.2.1 ``new Settings()`` + ``new World<>()`` gravity."""
    s = vsr_settings()
    s.step_frequency = 1.0 / 60.0
    s.velocity_iterations = 10
    s.position_iterations = 10
    s.linear_tolerance = 0.005
    s.angular_tolerance = math.radians(2.0)
    s.max_linear_correction = 0.2
    s.baumgarte = 0.2
    s.restitution_velocity = 1.0
    s.max_translation = 2.0
    s.max_rotation = 0.5 * math.pi
    s.gravity_x = 0.0
    s.gravity_y = -9.8
    s.ground_friction = 0.2
    s.ground_restitution = 0.0
    s.spring_mass_mode = VSR_SPRING_MASS_REDUCED
    s.self_collision = 1
    s.continuous_detection = 1  # ContinuousDetectionMode.ALL
    return s


def default_material() -> vsr_material:
    """``new Voxel(sensors)`` defaults, Voxel.java:57-68,144-160."""
    m = vsr_material()
    m.side_length = 3.0
    m.mass_side_length_ratio = 0.35
    m.spring_f = 8.0
    m.spring_d = 0.3
    m.mass_linear_damping = 0.1
    m.mass_angular_damping = 0.1
    m.friction = 10.0
    m.restitution = 0.1
    m.mass = 1.0
    m.area_ratio_passive_min = -math.inf
    m.area_ratio_passive_max = math.inf
    m.area_ratio_active_min = 0.8
    m.area_ratio_active_max = 1.2
    m.spring_scaffoldings = VSR_SCAFFOLD_ALL
    return m


_lib = None


def _declare(lib) -> None:
    vp, sz = C.c_void_p, C.c_size_t
    lib.vsr_default_settings.argtypes = [C.POINTER(vsr_settings)]
    lib.vsr_default_settings.restype = None
    lib.vsr_default_material.argtypes = [C.POINTER(vsr_material)]
    lib.vsr_default_material.restype = None
    lib.vsr_tick_count.argtypes = [C.c_double, C.c_double]
    lib.vsr_tick_count.restype = C.c_int
    lib.vsr_tick_count_from.argtypes = [C.c_double, C.c_double, C.c_double]
    lib.vsr_tick_count_from.restype = C.c_int
    lib.vsr_batch_create.argtypes = [C.POINTER(vsr_batch_desc), C.POINTER(vp)]
    lib.vsr_batch_create.restype = C.c_int
    lib.vsr_batch_upload.argtypes = [vp, C.c_int, vp]
    lib.vsr_batch_upload.restype = C.c_int
    lib.vsr_batch_run.argtypes = [vp, vp]
    lib.vsr_batch_run.restype = C.c_int
    lib.vsr_batch_results.argtypes = [vp, C.POINTER(vsr_result), sz]
    lib.vsr_batch_results.restype = C.c_int
    lib.vsr_batch_results_device.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    lib.vsr_batch_results_device.restype = C.c_int
    lib.vsr_batch_trajectory.argtypes = [vp, sz, C.POINTER(C.c_double), sz]
    lib.vsr_batch_trajectory.restype = C.c_longlong
    lib.vsr_batch_final_bbox.argtypes = [vp, C.POINTER(C.c_double), sz]
    lib.vsr_batch_final_bbox.restype = C.c_int
    lib.vsr_batch_voxel_trace.argtypes = [vp, sz, C.POINTER(C.c_double), sz]
    lib.vsr_batch_voxel_trace.restype = C.c_longlong
    lib.vsr_batch_sensor_readings.argtypes = [vp, sz, C.POINTER(C.c_double), sz]
    lib.vsr_batch_sensor_readings.restype = C.c_longlong
    lib.vsr_batch_voxel_steps.argtypes = [vp]
    lib.vsr_batch_voxel_steps.restype = sz
    lib.vsr_batch_n_ticks.argtypes = [vp]
    lib.vsr_batch_n_ticks.restype = C.c_int
    lib.vsr_batch_n_bodies.argtypes = [vp, sz]
    lib.vsr_batch_n_bodies.restype = C.c_int
    lib.vsr_batch_last_launches.argtypes = [vp]
    lib.vsr_batch_last_launches.restype = C.c_int
    lib.vsr_batch_upload_bytes.argtypes = [vp]
    lib.vsr_batch_upload_bytes.restype = sz
    lib.vsr_batch_destroy.argtypes = [vp]
    lib.vsr_batch_destroy.restype = None
    lib.vsr_last_error.argtypes = []
    lib.vsr_last_error.restype = C.c_char_p
    lib.vsr_abi_version.argtypes = []
    lib.vsr_abi_version.restype = C.c_int
    lib.vsr_device_count.argtypes = []
    lib.vsr_device_count.restype = C.c_int


EXPORTED_SYMBOLS = [
    "vsr_default_settings", "vsr_default_material", "vsr_tick_count", "vsr_tick_count_from", "vsr_batch_create",
    "vsr_batch_upload", "vsr_batch_run", "vsr_batch_results", "vsr_batch_results_device",
    "vsr_batch_trajectory", "vsr_batch_final_bbox", "vsr_batch_voxel_trace", "vsr_batch_sensor_readings", "vsr_batch_voxel_steps", "vsr_batch_n_ticks", "vsr_batch_n_bodies",
    "vsr_batch_last_launches", "vsr_batch_upload_bytes", "vsr_batch_destroy", "vsr_last_error", "vsr_abi_version",
    "vsr_device_count",
]


def load_library():
    """Load the CUDA product library.  Fails loudly: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        _declare(lib)
        if lib.vsr_abi_version() != VSR_ABI_VERSION:
            raise RuntimeError("libvsr_b200.so ABI version mismatch")
        _lib = lib
    return _lib
