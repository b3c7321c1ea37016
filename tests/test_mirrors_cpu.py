"""CPU: host-side mirrors of the reference's Python interface against golden vectors produced by
executing the reference (edge-list reader, pair sampler)."""
import os
import random

import numpy as np

from deepmerge_b200 import MyUtils1, MyUtils2
from oracle.ref_shim import FakeFeature, FakeLayer, FakeRaster


def test_edge_reader_matches_reference(golden_dir, capsys):
    e = np.load(os.path.join(golden_dir, "edge_reader.npz"))
    lines = FakeLayer([FakeFeature(i, {"LEFT_FID": int(l), "RIGHT_FID": int(r)}) for i, (l, r) in
                       enumerate(zip(e["left"], e["right"]))])
    ds = MyUtils2.PolygonConnectPointDataset("X:\\img\\tileA.tif", "X:\\tiles\\tileA.shp", "X:\\tiles\\tileA\\lines.shp",
                                             "X:\\tiles\\tileA\\PointsGCS.shp", line_layer=lines, polygon_layer=FakeLayer([]),
                                             point_layer=FakeLayer([]), img_dataset=FakeRaster(np.zeros((3, 8, 8), np.uint8)))
    assert len(ds) == len(e["out_fid"])
    rows = [ds[i] for i in range(len(ds))]
    assert [r[0] for r in rows] == e["out_fid"].tolist() and [r[1] for r in rows] == e["out_name"].tolist()
    assert [r[2] for r in rows] == e["out_left"].tolist() and [r[3] for r in rows] == e["out_right"].tolist()
    keys = ds.edge_keys()
    assert np.array_equal(keys >> 32, np.minimum(e["out_left"], e["out_right"]))
    assert ds.band_num == 3 and ds.line_layer is lines


def test_edge_reader_without_gdal_raises_like_the_reference():
    import pytest
    with pytest.raises(ValueError, match="Can not open"):
        MyUtils2.PolygonConnectPointDataset("a.tif", "a.shp", "lines.shp", "pts.shp")


def test_pair_sampler_reproduces_reference_under_the_same_seed(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "pair_sampler.npz"))
    fields = list(g["fields"])
    p = tmp_path / "tileP.txt"
    p.write_text("".join(f"{j},{a},{b},0,0\n" for j, (a, b) in enumerate(g["pos_pairs"])))
    pairs = MyUtils1.read_pair_list(str(p))
    assert [[int(a), int(b)] for a, b in pairs] == g["pos_pairs"].tolist()
    random.seed(int(g["seed"]))
    data = MyUtils1.sample_pairs(fields, pairs, 1, "tileP") + MyUtils1.sample_pairs(fields, g["neg_pairs"].tolist(), 0, "tileN")
    left, right, flag = MyUtils1.pairs_to_arrays(data)
    assert np.array_equal(left, g["out_left"]) and np.array_equal(right, g["out_right"]) and np.array_equal(flag, g["out_flag"])
    assert [d[0] for d in data] == g["out_tile"].tolist()
    assert g["counts"].tolist() == [len(g["pos_pairs"]), len(g["pos_pairs"]), len(g["neg_pairs"]), len(g["neg_pairs"])]
