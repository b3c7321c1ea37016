"""Builds the oracle's C restatement (oracle/dm_oracle.c -> oracle/_build/libdm_oracle.so) with gcc and loads it
with ctypes.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dm_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdm_oracle.so")


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        r = subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", LIB, SRC], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed on dm_oracle.c:\n" + r.stderr)
        if verbose:
            print("built", LIB)
    return LIB


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        p, i64, i = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.dmo_build_rag.restype = i64
        L.dmo_build_rag.argtypes = [p, i64, i64, i64, i64, i, i, p, p, p, p, i64]
        L.dmo_pool_bands.restype = None
        L.dmo_pool_bands.argtypes = [p, p, i64, i64, i64, i64, p, p]
        L.dmo_min_roots.restype = None
        L.dmo_min_roots.argtypes = [i64, p, p, i64, p]
        L.dmo_relabel.restype = None
        L.dmo_relabel.argtypes = [p, i64, p, i64, p]
        _LIB = L
    return _LIB


def build_rag(labels, n_regions, top_border=True, bottom_border=True, own_rows=None):
    """Same contract as oracle_np.build_rag -> (keys uint64, blen uint32, area int64, perim int64)."""
    import numpy as np
    L = np.ascontiguousarray(labels, np.int32)
    H, W = L.shape
    own = H if own_rows is None else own_rows
    cap = 2 * H * W + 1
    area, perim = np.zeros(n_regions, np.int64), np.zeros(n_regions, np.int64)
    keys, blen = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    e = lib().dmo_build_rag(L.ctypes.data, own, H, W, n_regions, int(top_border), int(bottom_border), area.ctypes.data,
                            perim.ctypes.data, keys.ctypes.data, blen.ctypes.data, cap)
    if e < 0:
        raise ValueError("dmo_build_rag failed (label >= n_regions?)")
    return keys[:e].copy(), blen[:e].copy(), area, perim


def pool_bands(labels, image, n_regions):
    import numpy as np
    L = np.ascontiguousarray(labels, np.int32)
    img = np.ascontiguousarray(image, np.uint8)
    C = img.shape[-1]
    s, q = np.zeros((n_regions, C), np.uint64), np.zeros((n_regions, C), np.uint64)
    lib().dmo_pool_bands(L.ctypes.data, img.ctypes.data, L.shape[0], L.shape[1], C, n_regions, s.ctypes.data, q.ctypes.data)
    return s, q


def min_roots(n, u, v):
    import numpy as np
    u, v = np.ascontiguousarray(u, np.int32), np.ascontiguousarray(v, np.int32)
    root = np.zeros(n, np.int32)
    lib().dmo_min_roots(n, u.ctypes.data, v.ctypes.data, len(u), root.ctypes.data)
    return root


def relabel(labels, root):
    import numpy as np
    L = np.ascontiguousarray(labels, np.int32)
    r = np.ascontiguousarray(root, np.int32)
    out = np.empty_like(L)
    lib().dmo_relabel(L.ctypes.data, L.size, r.ctypes.data, len(r), out.ctypes.data)
    return out
