"""Drop-in for the adjacent-pair sampler of the reference's MyUtils1.py (R10).

MergingSegmensPairDataset.add_data (MyUtils1.py:236-295): for every labelled polygon pair of a
pair-list txt (columns 1,2 = left,right polygon id, :225-234) pick ONE member sample point on each
side with two `random.randint` draws in that order (:278-279), and emit [tile, left_pt, right_pt,
flag].  The same draws in the same order reproduce the reference's list under the same seed.
`sample_pairs` is the array form used by the GPU training step (dm_gather_rows + Loss)."""
from __future__ import annotations

import random

import numpy as np


def read_pair_list(txt_path):
    """MyUtils1.py:225-234: comma separated lines, columns 1 and 2 are the polygon ids."""
    out = []
    with open(txt_path, "r") as f:
        for line in f.readlines():
            cols = line.strip("\n").split(",")
            out.append([cols[1], cols[2]])
    return out


def sample_pairs(point_id_fields, pairs, flag, tile="", rng=random):
    """-> list of [tile, left_point, right_point, flag] (point ids as the strings of the field)."""
    data = []
    for left_id, right_id in pairs:
        ls = point_id_fields[int(left_id)].split(" ")
        rs = point_id_fields[int(right_id)].split(" ")
        m_rand = rng.randint(0, len(ls) - 1)
        n_rand = rng.randint(0, len(rs) - 1)
        data.append([tile, ls[m_rand], rs[n_rand], flag])
    return data


def pairs_to_arrays(data):
    """-> (left int64 [B], right int64 [B], flag int64 [B]) for dm_gather_rows / Loss."""
    left = np.asarray([int(d[1]) for d in data], np.int64)
    right = np.asarray([int(d[2]) for d in data], np.int64)
    flag = np.asarray([int(d[3]) for d in data], np.int64)
    return left, right, flag
