"""CPU: the oracle's two independent restatements of the PARITY-UNPINNED integer stages (numpy in oracle_np.py,
plain C in dm_oracle.c) must agree on random, nodata-riddled and row-tiled rasters."""
import numpy as np
import pytest

from oracle import build as oc
from oracle import oracle_np as o


@pytest.mark.parametrize("seed", range(6))
def test_c_and_numpy_rag_agree(seed):
    rng = np.random.default_rng(seed)
    H, W, R = int(rng.integers(1, 60)), int(rng.integers(1, 90)), int(rng.integers(1, 40))
    small = rng.integers(0, R, size=((H + 4) // 5, (W + 5) // 6)).astype(np.int32)
    L = np.kron(small, np.ones((5, 6), np.int32))[:H, :W].copy()
    L[rng.random((H, W)) < 0.05] = rng.integers(0, R)
    L[rng.random((H, W)) < 0.03] = -1
    for a, b in zip(oc.build_rag(L, R), o.build_rag(L, R)):
        assert np.array_equal(a, b)
    if H > 3:                                             # a row tile with a halo row, interior borders
        own = H // 2
        tile = L[: own + 1]
        for a, b in zip(oc.build_rag(tile, R, top_border=True, bottom_border=False, own_rows=own),
                        o.build_rag(tile, R, top_border=True, bottom_border=False, own_rows=own)):
            assert np.array_equal(a, b)
    img = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8)
    for a, b in zip(oc.pool_bands(L, img, R), o.pool_bands(L, img, R)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(4))
def test_c_and_numpy_component_roots_agree(seed):
    rng = np.random.default_rng(100 + seed)
    n, m = 500, 420
    u, v = rng.integers(0, n, size=m), rng.integers(0, n, size=m)
    want = o.min_root_propagate(n, u, v)
    assert np.array_equal(oc.min_roots(n, u, v), want.astype(np.int32))
    L = rng.integers(-1, n, size=(30, 40)).astype(np.int32)
    assert np.array_equal(oc.relabel(L, want), o.relabel(L, want.astype(np.int32)))


def test_c_oracle_matches_synthetic_scene_merge():
    sc = o.synth_scene(96, 128, 160, C=4)
    res = o.merge_scene(sc["labels"], sc["n_regions"], sc["region_of_point"], sc["feats"], tau=0.5)
    k, b, area, per = oc.build_rag(sc["labels"], sc["n_regions"])
    assert np.array_equal(k, res["keys0"])
    k2, b2, area2, per2 = oc.build_rag(res["labels"], sc["n_regions"])   # a recount of the merged map
    roots = np.unique(res["root"])
    assert np.array_equal(k2, res["keys"]) and np.array_equal(b2, res["blen"])
    assert np.array_equal(area2[roots], res["area"][roots]) and np.array_equal(per2[roots], res["perim"][roots])
