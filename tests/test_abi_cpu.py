"""The C-ABI library: loads, exports every declared symbol, validates like the reference."""
import ctypes as C
import importlib
import math
import os
import re

import numpy as np
import pytest

H = importlib.import_module("2dhmsr_b200")
A = H.abi
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vsr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vsr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(A.EXPORTED_SYMBOLS)
    lib = C.CDLL(A.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert A.load_library().vsr_abi_version() == A.VSR_ABI_VERSION


def test_struct_layouts_match_the_header():
    # sizes implied by the header (x86-64 SysV): guards against a silent ctypes drift
    assert C.sizeof(A.vsr_settings) == 8 + 4 + 4 + 11 * 8 + 8 + 8
    assert C.sizeof(A.vsr_material) == 13 * 8 + 8
    assert C.sizeof(A.vsr_sensor) == 6 * 4 + 4 * 8 + 8 * 8
    assert C.sizeof(A.vsr_result) == 88 and H.RESULT_DTYPE.itemsize == 88
    lib = A.load_library()
    s, m = A.vsr_settings(), A.vsr_material()
    lib.vsr_default_settings(C.byref(s))
    lib.vsr_default_material(C.byref(m))
    assert bytes(s) == bytes(A.default_settings()) and bytes(m) == bytes(A.default_material())
    assert s.step_frequency == 1 / 60 and s.velocity_iterations == 10 and s.gravity_y == -9.8
    assert m.side_length == 3.0 and m.spring_f == 8.0 and math.isinf(m.area_ratio_passive_max)


def _create(bd):
    lib = A.load_library()
    h = C.c_void_p()
    rc = lib.vsr_batch_create(bd.ptr(), C.byref(h))
    msg = lib.vsr_last_error().decode()
    if rc == 0:
        lib.vsr_batch_destroy(h)
    return rc, msg


def test_create_validates_like_the_reference():
    loco = helpers.locomotion(2.0)
    ok = loco.compile(helpers.phase_robots(2))
    assert _create(ok)[0] == 0
    bd = loco.compile(helpers.phase_robots(2))
    bd.desc.abi_version = 99
    assert _create(bd)[0] == A.VSR_ERR_INVALID_ARGUMENT
    bd = loco.compile(helpers.phase_robots(2))
    xs = np.asarray([0.0, 50.0, 20.0, 100.0])
    bd.desc.terrain_xs = xs.ctypes.data_as(C.POINTER(C.c_double))
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "x coordinates must be sorted" in msg  # Ground.java:55-58
    bd = loco.compile(helpers.phase_robots(2))
    bd.desc.terrain_n = 1
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "at least 2 points" in msg  # Ground.java:52-54
    bd = loco.compile(helpers.phase_robots(2))
    bd.desc.material.area_ratio_passive_min = 0.2
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "areaRatioPassiveRange" in msg  # Voxel.java:132-140
    bd = loco.compile(helpers.phase_robots(2))
    bd.desc.material.area_ratio_passive_min, bd.desc.material.area_ratio_passive_max = 0.7, 1.3
    assert _create(bd)[0] == A.VSR_OK                                          # rope limits: on the accelerated path
    bd = loco.compile(helpers.phase_robots(2))
    bd.desc.material.area_ratio_passive_min, bd.desc.material.area_ratio_passive_max = 1.05, 1.3
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "Wrong spring range" in msg    # SpringRange, Voxel.java:189-193
    bd = loco.compile(helpers.phase_robots(2))
    brk = (A.vsr_breakable * 10)()
    brk[3].enabled = 1
    brk[3].malfunctions[A.VSR_COMPONENT_STRUCTURE] = 1 << A.VSR_MALFUNCTION_RANDOM
    bd.desc.shapes[0].breakable = C.cast(brk, C.POINTER(A.vsr_breakable))
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "Unsupported structure malfunction type" in msg  # BreakableVoxel.java:258
    bd = loco.compile(helpers.phase_robots(2))
    offs = np.asarray([0, 12, 23], dtype=np.uint64)
    bd.desc.param_offsets = C.cast(offs.ctypes.data, C.POINTER(C.c_size_t))
    assert _create(bd)[0] == A.VSR_ERR_INVALID_ARGUMENT
    bd = loco.compile(helpers.mlp_robots(2, hidden=(5,)))
    offs = np.asarray([0, 10, 20], dtype=np.uint64)
    bd.desc.param_offsets = C.cast(offs.ctypes.data, C.POINTER(C.c_size_t))
    rc, msg = _create(bd)
    assert rc == A.VSR_ERR_INVALID_ARGUMENT and "Wrong number of weights" in msg  # MultiLayerPerceptron.java:56-62
    assert _create(helpers.locomotion(1.0).compile(helpers.phase_robots(1, shape="box-10x10")))[0] == 0
    big = helpers.locomotion(1.0).compile(helpers.phase_robots(1, shape="box-12x12"))
    rc, msg = _create(big)
    assert rc == A.VSR_ERR_UNSUPPORTED and "128 voxels" in msg
    with pytest.raises(ValueError):
        H.DeviceBatch(bd)
    with pytest.raises(NotImplementedError):
        H.DeviceBatch(big)


def test_batch_facts_without_a_device():
    lib = A.load_library()
    bd = helpers.locomotion().compile(helpers.phase_robots(3) + helpers.phase_robots(1, shape="worm-8x2"))
    h = C.c_void_p()
    assert lib.vsr_batch_create(bd.ptr(), C.byref(h)) == 0
    try:
        assert lib.vsr_batch_n_ticks(h) == 1200
        assert lib.vsr_batch_voxel_steps(h) == bd.voxel_steps() == (30 + 16) * 1200
        assert lib.vsr_batch_n_bodies(h, 0) == 40 and lib.vsr_batch_n_bodies(h, 3) == 64
        out = (A.vsr_result * 4)()
        assert lib.vsr_batch_results(h, out, 4) == A.VSR_ERR_STATE  # results before run
        assert lib.vsr_batch_run(h, None) == A.VSR_ERR_STATE         # run before upload
        if lib.vsr_device_count() == 0:
            # no CPU fallback: the compute entry points refuse loudly
            assert lib.vsr_batch_upload(h, 0, None) == A.VSR_ERR_NO_DEVICE
            assert b"no CPU fallback" in lib.vsr_last_error()
    finally:
        lib.vsr_batch_destroy(h)


def test_product_path_never_touches_the_oracle():
    """The product sources must not reference oracle/ (a CPU fallback would void parity)."""
    pk = os.path.join(ROOT, "2dhmsr_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                for pat in (r"libvsr_oracle", r"vsr_oracle_\w+\s*\(", r"^\s*(from|import)\s+oracle", r"oracle/",
                            r"#include.*oracle", r"oracle\.(run|lib|build)"):
                    assert not re.search(pat, src, flags=re.M), (pat, os.path.join(dp, f))


def test_create_rejects_a_ground_profile_finer_than_the_contact_capacity():
    """The device keeps 3 ground constraints per mass (enough for every terrain Locomotion.createTerrain builds:
    segments >= 1 wide, steppy risers 0.5 wide between treads >= 1 wide).  A profile on which a mass could touch
    more polygons at once is refused at create -- never truncated at run time."""
    import helpers
    robots = helpers.phase_robots(1)
    for name in ("flat", "hilly-1-10-0", "hilly-0.5-1-3", "steppy-1-6-2", "steppy-2-1-9", "uphill-10", "flatWithStart-4"):
        H.DeviceBatch(helpers.locomotion(1.0, terrain=name).compile(robots)).close()   # vsr_batch_create (host only) accepts it
    xs = [0.0, 10.0] + [10.0 + 0.3 * i for i in range(1, 200)] + [100.0]
    ys = [100.0, 5.0] + [5.0 + 0.05 * (i % 2) for i in range(1, 200)] + [100.0]
    with pytest.raises(NotImplementedError, match="ground profile too fine"):
        H.DeviceBatch(H.Locomotion(1.0, [xs, ys], H.Settings()).compile(robots))


def test_create_rejects_more_ground_points_than_the_time_of_impact_pass_can_index():
    """The pass packs a ground polygon's index into 17 bits of its candidate list: a longer profile is refused when
    continuous detection is on (the default), accepted without it."""
    import helpers
    import numpy as np
    robots = helpers.phase_robots(1)
    n = (1 << 17) + 2
    xs, ys = list(2.0 * np.arange(n)), [5.0] * n
    H.DeviceBatch(H.Locomotion(1.0, [xs, ys], H.Settings().set(continuous_detection=0)).compile(robots)).close()
    with pytest.raises(NotImplementedError, match="continuous detection: more than 131072 ground points"):
        H.DeviceBatch(H.Locomotion(1.0, [xs, ys], H.Settings()).compile(robots))
