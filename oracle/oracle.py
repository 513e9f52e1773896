"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE -- see vsr_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libvsr_oracle.so")

ORDER_CANONICAL, ORDER_INSERTION, ORDER_SHUFFLED, ORDER_DFS = 0, 1, 2, 3


class vsr_oracle_opts(C.Structure):
    _fields_ = [
        ("order", C.c_int32), ("seed", C.c_uint32), ("n_threads", C.c_int32), ("max_ticks", C.c_int32),
        ("robot_begin", C.c_size_t), ("robot_end", C.c_size_t), ("self_collision", C.c_int32),
        ("ccd", C.c_int32), ("state_fp32", C.c_int32), ("sleeping", C.c_int32), ("dyn4j_labels", C.c_int32), ("touching_separated", C.c_int32),
    ]


class vsr_oracle_probe(C.Structure):
    _fields_ = [
        ("trajectory", C.POINTER(C.c_double)), ("body_stride", C.c_size_t),
        ("signals", C.POINTER(C.c_double)), ("voxel_stride", C.c_size_t),
        ("readings", C.POINTER(C.c_double)), ("reading_stride", C.c_size_t),
        ("n_contacts", C.POINTER(C.c_int32)), ("pos_iters", C.POINTER(C.c_int32)),
        ("touch", C.POINTER(C.c_double)),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "vsr_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B" if force else "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        pkg = importlib.import_module("2dhmsr_b200")
        A = pkg.abi
        l.vsr_oracle_run.argtypes = [C.POINTER(A.vsr_batch_desc), C.POINTER(vsr_oracle_opts),
                                     C.POINTER(A.vsr_result), C.POINTER(vsr_oracle_probe)]
        l.vsr_oracle_run.restype = C.c_int
        l.vsr_oracle_mlp_apply.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
        l.vsr_oracle_mlp_apply.restype = C.c_int
        l.vsr_oracle_window_first.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int32), C.c_int]
        l.vsr_oracle_window_first.restype = C.c_int
        l.vsr_oracle_tick_times.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int]
        l.vsr_oracle_tick_times.restype = C.c_int
        l.vsr_oracle_last_error.restype = C.c_char_p
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run(bd, order=ORDER_CANONICAL, seed=0, n_threads=1, max_ticks=0, robot_begin=0, robot_end=0,
        self_collision=False, trajectory=False, signals=False, readings=False, stats=False, ccd=False, state_fp32=False, sleeping=False, dyn4j_labels=False, touching_separated=False):
    """Run the oracle on a BatchDesc.  Returns (results, probes dict)."""
    pkg = importlib.import_module("2dhmsr_b200")
    l = lib()
    n = int(bd.desc.n_robots)
    end = n if robot_end in (0, None) else min(robot_end, n)
    rows = end - robot_begin
    out = np.zeros(rows, dtype=pkg.RESULT_DTYPE)
    o = vsr_oracle_opts(order, seed, n_threads, max_ticks, robot_begin, end, 1 if self_collision else 0, 1 if ccd else 0,
                        1 if state_fp32 else 0, 1 if sleeping else 0, 1 if dyn4j_labels else 0, 1 if touching_separated else 0)
    p = vsr_oracle_probe()
    probes = {}
    nt = bd.n_ticks
    maxv = max(bd.n_voxels)
    if trajectory:
        probes["trajectory"] = np.zeros((rows, nt, 4 * maxv, 6))
        p.trajectory, p.body_stride = _dp(probes["trajectory"]), 4 * maxv
    if signals:
        probes["signals"] = np.zeros((rows, nt, maxv))
        probes["touch"] = np.zeros((rows, nt, maxv))
        p.signals, p.voxel_stride = _dp(probes["signals"]), maxv
        p.touch = _dp(probes["touch"])
    if readings:
        mr = max(1, max(bd.n_readings))
        probes["readings"] = np.zeros((rows, nt, mr))
        p.readings, p.reading_stride = _dp(probes["readings"]), mr
    if stats:
        probes["n_contacts"] = np.zeros((rows, nt), dtype=np.int32)
        probes["pos_iters"] = np.zeros((rows, nt), dtype=np.int32)
        p.n_contacts = probes["n_contacts"].ctypes.data_as(C.POINTER(C.c_int32))
        p.pos_iters = probes["pos_iters"].ctypes.data_as(C.POINTER(C.c_int32))
    rc = l.vsr_oracle_run(bd.ptr(), C.byref(o), C.cast(out.ctypes.data, C.POINTER(pkg.abi.vsr_result)), C.byref(p))
    if rc != 0:
        msg = l.vsr_oracle_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        if rc == -2:
            raise NotImplementedError(msg)
        raise RuntimeError(f"oracle error {rc}: {msg}")
    return out, probes


def mlp_apply(activation: int, neurons, weights, x):
    l = lib()
    ne = np.asarray(neurons, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    xi = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(int(ne[-1]))
    l.vsr_oracle_mlp_apply(activation, ne.ctypes.data_as(C.POINTER(C.c_int32)), len(ne), _dp(w), _dp(xi), _dp(out))
    return out


def window_first(final_t, dt, interval):
    l = lib()
    nt = l.vsr_oracle_tick_times(final_t, dt, None, 0)
    f = np.zeros(nt, dtype=np.int32)
    l.vsr_oracle_window_first(final_t, dt, interval, f.ctypes.data_as(C.POINTER(C.c_int32)), nt)
    return f


def tick_times(final_t, dt):
    l = lib()
    nt = l.vsr_oracle_tick_times(final_t, dt, None, 0)
    t = np.zeros(nt)
    l.vsr_oracle_tick_times(final_t, dt, _dp(t), nt)
    return t
