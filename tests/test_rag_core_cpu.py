"""CPU: the per-block arithmetic of the raster kernel.  deepmerge_b200/csrc/rag_core.cuh holds the functions the CUDA
kernel calls for every 4 x 4 block (window test, fast-path statistics, process_item with all the border / nodata /
pair rules); the same header is compiled here with g++ (tests/rag_core_host.cpp walks the raster lane by lane the way
the kernel does) and checked bit for bit against the oracle.  What is left to the GPU tests is the staging, the work
list and the hash tables around these functions."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_np as o

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rag_core") / "librag_core_host.so")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "rag_core_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(so)
    lib.rag_core_host.restype = ctypes.c_long
    lib.rag_core_host.argtypes = [ctypes.c_void_p] + [ctypes.c_long] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_long,
                                                                             ctypes.c_long, ctypes.c_int, ctypes.c_int] + \
        [ctypes.c_void_p] * 6 + [ctypes.c_long, ctypes.c_void_p]

    def run(L, R, image=None, own=None, top=True, bottom=True):
        L = np.ascontiguousarray(L, np.int32)
        H, W = L.shape
        own = H if own is None else own
        C = 0 if image is None else image.shape[2]
        img = None if image is None else np.ascontiguousarray(image[:own], np.uint8)
        area = np.zeros(R, np.int64)
        border = np.zeros(R, np.int64)
        bs = np.zeros((R, max(C, 1)), np.uint64)
        bq = np.zeros((R, max(C, 1)), np.uint64)
        cap = 2 * H * W + 8
        keys = np.zeros(cap, np.uint64)
        cnt = np.zeros(cap, np.uint32)
        stats = np.zeros(2, np.int64)
        n = lib.rag_core_host(L.ctypes.data, own, H, W, W, img.ctypes.data if img is not None else None, C, W * C, R,
                              int(top), int(bottom), area.ctypes.data, border.ctypes.data, bs.ctypes.data, bq.ctypes.data,
                              keys.ctypes.data, cnt.ctypes.data, cap, stats.ctypes.data)
        assert n >= 0, n
        return dict(keys=keys[:n], blen=cnt[:n], area=area, border=border, bsum=bs[:, :C], bsq=bq[:, :C], items=int(stats[0]),
                    fast=int(stats[1]))
    return run


def check(host, L, R, image=None, own=None, top=True, bottom=True):
    got = host(L, R, image, own, top, bottom)
    keys, blen, area, per = o.build_rag(L, R, top_border=top, bottom_border=bottom, own_rows=own)
    assert np.array_equal(got["keys"], keys)
    assert np.array_equal(got["blen"], blen)
    assert np.array_equal(got["area"], area)
    lo, hi = o.unpack_keys(got["keys"])
    per_got = got["border"].copy()
    np.add.at(per_got, lo.astype(np.int64), got["blen"].astype(np.int64))
    np.add.at(per_got, hi.astype(np.int64), got["blen"].astype(np.int64))
    assert np.array_equal(per_got, per)
    if image is not None:
        n = L.shape[0] if own is None else own
        s, q = o.pool_bands(L[:n], image[:n], R)
        assert np.array_equal(got["bsum"], s)
        assert np.array_equal(got["bsq"], q)
    return got


@pytest.mark.parametrize("H,W", [(1, 1), (1, 7), (5, 1), (4, 4), (33, 128), (32, 256), (33, 257), (64, 260), (65, 516), (100, 300)])
@pytest.mark.parametrize("C", [0, 4])
def test_blocky_rasters_with_speckle(host, H, W, C):
    rng = np.random.default_rng(H * 1000 + W + C)
    R = 37
    small = rng.integers(0, R, size=((H + 5) // 6, (W + 6) // 7)).astype(np.int32)
    L = np.kron(small, np.ones((6, 7), np.int32))[:H, :W].copy()
    L[rng.random((H, W)) < 0.02] = rng.integers(0, R)
    img = rng.integers(0, 256, size=(H, W, C)).astype(np.uint8) if C else None
    check(host, L, R, img)


@pytest.mark.parametrize("C", [1, 2, 3, 4])
def test_synthetic_scene_all_band_counts(host, C):
    sc = o.synth_scene(150, 512, 300, C=C)
    check(host, sc["labels"], sc["n_regions"], sc["image"])


def test_most_blocks_of_a_bench_like_scene_take_the_fast_path(host):
    sc = o.synth_scene(512, 1024, 512, C=4)             # ~1000-pixel regions, as in BASELINE configs[1]
    got = check(host, sc["labels"], sc["n_regions"], sc["image"])
    assert got["fast"] > 2 * got["items"], (got["fast"], got["items"])


def test_noise_nodata_and_one_pixel_regions(host):
    rng = np.random.default_rng(5)
    H, W = 70, 300
    L = rng.integers(0, 50, size=(H, W)).astype(np.int32)
    L[rng.random((H, W)) < 0.1] = -1
    L[10:20, 40:90] = -1
    L[30:33, 100:140] = -7                              # a second nodata value
    img = rng.integers(0, 256, size=(H, W, 4)).astype(np.uint8)
    check(host, L, 50, img)
    check(host, np.arange(H * W, dtype=np.int32).reshape(H, W), H * W, img)
    check(host, np.full((H, W), -1, np.int32), 5, img)
    check(host, np.zeros((H, W), np.int32), 1, img)


def test_two_and_three_label_windows_with_nodata(host):
    rng = np.random.default_rng(11)
    for trial in range(40):
        H, W = int(rng.integers(6, 40)), int(rng.integers(6, 200))
        k = int(rng.integers(2, 5))
        small = rng.integers(-1, k, size=((H + 2) // 3, (W + 4) // 5)).astype(np.int32)
        L = np.kron(small, np.ones((3, 5), np.int32))[:H, :W].copy()
        img = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8)
        check(host, L, k, img)


@pytest.mark.parametrize("own,halo", [(64, 1), (65, 1), (1, 1), (3, 1), (4, 1), (7, 0), (8, 1), (105, 1)])
def test_row_tiles(host, own, halo):
    sc = o.synth_scene(110, 384, 400, C=4)
    L, R, img = sc["labels"], sc["n_regions"], sc["image"]
    for y0 in (0, 5):
        tile = L[y0:y0 + own + halo]
        if tile.shape[0] < own + halo:
            continue
        check(host, tile, R, img[y0:y0 + own + halo], own=own, top=(y0 == 0), bottom=(halo == 0))


def test_tiles_sum_to_the_whole(host):
    sc = o.synth_scene(90, 300, 200, C=2)
    L, R, img = sc["labels"], sc["n_regions"], sc["image"]
    whole = check(host, L, R, img)
    cuts = [0, 3, 4, 41, 90]
    area = np.zeros(R, np.int64)
    border = np.zeros(R, np.int64)
    edges = {}
    for i in range(len(cuts) - 1):
        y0, y1 = cuts[i], cuts[i + 1]
        last = y1 == 90
        g = host(L[y0:y1 + (0 if last else 1)], R, img[y0:y1 + (0 if last else 1)], own=y1 - y0, top=(y0 == 0), bottom=last)
        area += g["area"]
        border += g["border"]
        for k, n in zip(g["keys"].tolist(), g["blen"].tolist()):
            edges[k] = edges.get(k, 0) + n
    assert np.array_equal(area, whole["area"]) and np.array_equal(border, whole["border"])
    assert sorted(edges.items()) == list(zip(whole["keys"].tolist(), whole["blen"].tolist()))
