"""The plain-C example of the drop-in boundary (examples/c_abi_example.c): builds against include/vsr_b200.h
and the in-tree library; on a GPU it must give what the Python mirror gives for the same robots."""
import importlib
import json
import os
import subprocess

import pytest

H = importlib.import_module("2dhmsr_b200")
import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "c_abi_example")


def _build():
    lib_dir = os.path.join(ROOT, "2dhmsr_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", lib_dir, "-lvsr_b200",
                           f"-Wl,-rpath,{lib_dir}", "-lm", "-o", EXE])


def test_c_example_builds_against_the_header():
    H.abi.load_library()   # builds the library if needed
    _build()
    assert os.access(EXE, os.X_OK)


@pytest.mark.gpu
def test_c_example_equals_python_path():
    _build()
    j = json.loads(subprocess.check_output([EXE, "3"], text=True))
    body = helpers.body_of("biped-4x3")
    robots = [H.Robot(H.PhaseSin(1.0, 1.0, H.Grid.create(4, 3, lambda x, y, r=r: 0.1 * (r + 1) + 0.7 * x + 0.3 * y)), body)
              for r in range(3)]
    outs = H.Locomotion(20.0, H.Locomotion.create_terrain("flat"), H.Settings()).apply(robots)   # vsr_default_settings
    assert j["n_ticks"] == 1200
    assert j["distance"] == [o.get_distance() for o in outs]
    assert j["velocity"] == [o.get_velocity() for o in outs]
