"""Drop-in callables for the scoring half of the reference's ExtractFeatures.py.

Same names, argument meaning and return types as the reference (numpy in, numpy out), backed
by the CUDA library; there is no numpy fallback -- a CUDA device is required.

  Euclidean_distance(X, Y) -> D      ExtractFeatures.py:119-147
  MC_Lyu_2020(X, Y) -> D             ExtractFeatures.py:228-237 (same formula)
  pool_and_score(store, point_id_fields, left_ids, right_ids) -> (means, simi)
                                     the loop body of test_for_shp, ExtractFeatures.py:164-222,
                                     for ALL edges at once (the reference `break`s after the first)
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib
from .raster import _p, _stream, pool_points_csr, region_mean, score_l2


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("deepmerge_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def Euclidean_distance(X, Y):
    """X n*p, Y m*p -> D n*m with D[i,j] = sqrt(max(0, |X_i|^2 + |Y_j|^2 - 2 X_i.Y_j)), float32."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    if X.ndim != 2 or Y.ndim != 2 or X.shape[1] != Y.shape[1]:
        raise ValueError("X and Y must be [n,p] and [m,p]")
    dev = _dev()
    L = lib()
    x, y = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    out = torch.empty((X.shape[0], Y.shape[0]), dtype=torch.float32, device=dev)
    L.check(L.dm_euclidean_matrix(_p(x), _p(y), X.shape[0], Y.shape[0], X.shape[1], _p(out), _stream()), "dm_euclidean_matrix")
    return out.cpu().numpy()


def MC_Lyu_2020(X, Y):
    return Euclidean_distance(X, Y)


def membership_csr(point_id_fields, sep=" "):
    """'PointID' strings (ExtractFeatures.py:175-179) -> CSR (offsets int64 [R+1], ids int32 [N])."""
    offsets = np.zeros(len(point_id_fields) + 1, np.int64)
    ids = []
    for r, f in enumerate(point_id_fields):
        if f != "":
            ids.extend(int(t) for t in f.split(sep))
        offsets[r + 1] = len(ids)
    return offsets, np.asarray(ids, np.int32)


def pool_and_score(store, point_id_fields, left_ids, right_ids):
    """Mean-pool every polygon's member rows of the feature store (np.mean(axis=0) semantics, bit
    exact) and score every (left, right) edge with the Euclidean distance -> (means [R,D] float32,
    simi [E] float64, what the reference writes to the 'simi' OFTReal field :217-219)."""
    dev = _dev()
    off, ids = membership_csr(point_id_fields)
    store_t = torch.from_numpy(np.ascontiguousarray(store, dtype=np.float32)).to(dev)
    s, c = pool_points_csr(torch.from_numpy(off).to(dev), torch.from_numpy(ids).to(dev), store_t)
    mean, n2 = region_mean(s, c)
    left = torch.as_tensor(np.asarray(left_ids, np.int64), device=dev)
    right = torch.as_tensor(np.asarray(right_ids, np.int64), device=dev)
    keys = (torch.minimum(left, right) << 32) | torch.maximum(left, right)
    simi = score_l2(mean, keys, n2)
    return mean.cpu().numpy(), simi.cpu().numpy().astype(np.float64)
