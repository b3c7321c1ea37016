"""CPU: the arithmetic of dm_resize_area.  deepmerge_b200/csrc/resize_core.cuh holds the per-value functions the CUDA
kernel calls; the same header is compiled here with g++ (tests/resize_core_host.cpp runs the kernel's loops
sequentially) and checked bit for bit against the oracle restatement of cv2's INTER_AREA and against the golden
outputs of the executed reference -- with the tables the product builds (MyUtils2.area_tables).  What is left to the
GPU tests is the CUDA indexing around these functions."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from deepmerge_b200 import MyUtils2
from oracle.resize_area import resize_data

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("resize_core") / "libresize_core_host.so")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                        os.path.join(HERE, "resize_core_host.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    lib.resize_planes_host.restype = ctypes.c_int
    lib.resize_planes_host.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def run(planes, t):
        planes = np.ascontiguousarray(planes, np.uint8)
        n, s, _ = planes.shape
        mode, ti, tf = MyUtils2.area_tables(s, t)
        out = np.full((n, t, t), -1.0, np.float32)
        lib.resize_planes_host(planes.ctypes.data, n, s, t, mode, ti.ctypes.data if ti is not None else None,
                               tf.ctypes.data if tf is not None else None, out.ctypes.data)
        return out
    return run


def test_core_matches_the_executed_reference(host, golden_dir):
    g = np.load(os.path.join(golden_dir, "resize.npz"))
    for i, (s, t) in enumerate(g["cases"]):
        assert np.array_equal(host(g[f"in{i}"], int(t)), g[f"out{i}"]), (int(s), int(t))


def test_core_size_sweep_against_the_oracle(host):
    rng = np.random.default_rng(12)
    for t in (32, 64, 128, 1):
        for s in sorted(set(rng.integers(1, 3 * max(t, 40), 40).tolist() + [t, 2 * t, 3 * t, 4 * t, 5 * t, t + 1, max(t - 1, 1)])):
            x = rng.integers(0, 256, (2, s, s), dtype=np.uint8)
            if s > 3:
                x[0, : s // 2] = 0                                         # zero padding of a border window
                x[1, :, s // 3:] = 255                                     # saturation
            assert np.array_equal(host(x, t), resize_data(x, t)), (s, t)
