"""The second adjacency form of the reference: the polygon attribute `join`,
a comma-separated list of neighbour ids INCLUDING the polygon itself, which
RemoteSensingDataset.GenerateTestDataset removes (MyUtils.py:110-114); and the comma-separated `Points`
membership field (:116-117).  Host-side parsing into the packed edge keys / CSR the GPU path takes."""
from __future__ import annotations

import numpy as np


def neighbours_from_join(join_field, self_id):
    """'3,7,12' with self_id 7 -> [3, 12] (list.remove semantics: the first occurrence of self goes)."""
    ids = [int(t) for t in str(join_field).split(",")]
    ids.remove(int(self_id))
    return ids


def edge_keys_from_join(join_fields):
    """join fields of polygons 0..R-1 -> sorted unique packed (min << 32) | max keys (int64), the same
    canonical form build_rag produces from a label raster."""
    lo, hi = [], []
    for r, f in enumerate(join_fields):
        for n in neighbours_from_join(f, r):
            if n != r:
                lo.append(min(r, n))
                hi.append(max(r, n))
    keys = (np.asarray(lo, np.int64) << 32) | np.asarray(hi, np.int64)
    return np.unique(keys)


def membership_from_points(points_fields, sep=","):
    """'Points' fields (MyUtils.py:116-117) -> CSR (offsets int64 [R+1], ids int32 [N])."""
    offsets = np.zeros(len(points_fields) + 1, np.int64)
    ids = []
    for r, f in enumerate(points_fields):
        if str(f) != "":
            ids.extend(int(t) for t in str(f).split(sep))
        offsets[r + 1] = len(ids)
    return offsets, np.asarray(ids, np.int32)
