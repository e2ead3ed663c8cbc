"""The C++ host mirror (include/ccake_b200_host.hpp): compiles against the C ABI here (no GPU needed), and on the
GPU reproduces the Python mirror's run bit for bit (both are thin wrappers over the same library)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from common import ROOT, gubser_ic, viscous_settings

SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    from ccake_b200 import _capi

    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    deps = [SRC, os.path.join(ROOT, "include", "ccake_b200_host.hpp"), os.path.join(ROOT, "include", "ccake_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        libdir = os.path.dirname(_capi.LIB_PATH)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                               "-L", libdir, "-lccake_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_cpp_host_mirror_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_host_mirror_matches_python_mirror(tmp_path):
    from ccake_b200 import SPHWorkstation, to_config

    exe = _build()
    settings = viscous_settings()
    ic = gubser_ic(dx=0.1, shear=True)
    n, steps, dt = ic["r"].shape[0], 3, settings["dt"]
    cfg = to_config(settings)
    r3 = np.zeros((n, 3)); r3[:, :2] = ic["r"]
    u3 = np.zeros((n, 3)); u3[:, :2] = ic["u"]
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(bytes(cfg))
        f.write(struct.pack("<qqd", n, steps, dt))
        for a in (r3, ic["e"], u3, ic["shv6"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    subprocess.check_call([exe, str(inp), str(out)])
    raw = open(out, "rb").read()
    t, S, w = struct.unpack("<ddq", raw[:24])
    state_cpp = np.frombuffer(raw[24:], dtype=np.float64).reshape(w, n)

    ws = SPHWorkstation(settings)
    ws.upload_initial_conditions(ic["r"], ic["e"], u=ic["u"], shv6=ic["shv6"])
    for _ in range(steps):
        ws.advance_timestep(dt)
    ws.synchronize()
    state_py = np.empty((ws.state_width, n))
    ws.copy_device_to_host(state_py)
    assert w == ws.state_width and abs(t - ws.t) < 1e-15
    assert np.array_equal(state_cpp, state_py)
    assert abs(S - ws.conservation()["S"]) <= 1e-12 * abs(S)
