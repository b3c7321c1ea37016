"""CPU: the GDAL-free shapefile adaptor.  The reader is checked against
.dbf / .shp bytes laid out by hand from the published formats (independent of the module's own
writer), against the golden edge list produced by executing the reference, and through the
reference-shaped entry points (PolygonConnectPointDataset, score_layers = test_for_shp's loop)."""
import os
import struct

import numpy as np
import pytest

from deepmerge_b200 import MyUtils2, shapefile
from deepmerge_b200.ExtractFeatures import FeatureIO, membership_csr, score_layers


def raw_dbf(fields, rows, deleted=()):
    """dBase III bytes from the format description: 32-byte header, 32-byte field descriptors, 0x0D,
    fixed-width records with a leading deletion flag, 0x1A."""
    rlen = 1 + sum(w for _, _, w, _ in fields)
    hlen = 32 + 32 * len(fields) + 1
    out = bytearray(struct.pack("<BBBBIHH20x", 3, 124, 10, 18, len(rows), hlen, rlen))
    for name, t, w, d in fields:
        out += name.encode().ljust(11, b"\0") + t.encode() + b"\0" * 4 + bytes([w, d]) + b"\0" * 14
    out += b"\x0d"
    for i, row in enumerate(rows):
        out += b"*" if i in deleted else b" "
        for (name, t, w, d), cell in zip(fields, row):
            out += cell.ljust(w) if t == "C" else cell.rjust(w)
    return bytes(out + b"\x1a")


FIELDS = [("LEFT_FID", "N", 9, 0), ("RIGHT_FID", "N", 9, 0), ("PointID", "C", 30, 0), ("simi", "N", 24, 15),
          ("len", "F", 13, 3)]


def test_reader_against_hand_built_dbf(tmp_path):
    rows = [(b"3", b"7", b"10 11 12", b"0.250000000000000", b"1.5"),
            (b"-1", b"4", b"", b"", b"2.25"),
            (b"5", b"-1", b"99", b"************************", b"0.000"),
            (b"12", b"13", b"1 2", b"1.000000000000000", b"7.125")]
    p = tmp_path / "lines.dbf"
    p.write_bytes(raw_dbf(FIELDS, rows, deleted={3}))
    t = shapefile.DbfTable(str(p))
    assert [f.name for f in t.fields] == [f[0] for f in FIELDS] and len(t) == 4
    assert t.column_int("LEFT_FID").tolist() == [3, -1, 5, 12] and t.column_int("right_fid").tolist() == [7, 4, -1, 13]
    assert t.column_str("PointID") == ["10 11 12", "", "99", "1 2"]
    s = t.column_float("simi")
    assert s[0] == 0.25 and np.isnan(s[1]) and np.isnan(s[2]) and s[3] == 1.0
    assert t.deleted.tolist() == [False, False, False, True]
    # OGR typing of single cells: int for N(w,0), float otherwise, stripped str, None for null
    assert t.value(0, "LEFT_FID") == 3 and isinstance(t.value(0, "LEFT_FID"), int)
    assert t.value(0, "len") == 1.5 and isinstance(t.value(0, "len"), float)
    assert t.value(1, "simi") is None and t.value(2, "simi") is None and t.value(1, "PointID") == ""
    # the OGR-shaped layer: FID = record index, deleted records are skipped, missing FIDs give None
    layer = shapefile.ShapefileLayer(str(tmp_path / "lines.shp"))
    feats = list(layer)
    assert [f.GetFID() for f in feats] == [0, 1, 2] and layer.GetFeatureCount() == 3
    assert layer.GetFeature(3) is None and layer.GetFeature(17) is None
    assert feats[0].GetField("PointID").split(" ") == ["10", "11", "12"] and feats[0].GetField(0) == 3
    assert layer.GetLayerDefn().GetFieldIndex("simi") == 3 and layer.GetLayerDefn().GetFieldIndex("nope") == -1
    with pytest.raises(KeyError):
        feats[0].GetField("nope")


def test_write_read_round_trip_and_vectorised_columns(tmp_path):
    rng = np.random.default_rng(3)
    n = 5000
    a = rng.integers(-1, 10 ** 6, n)
    x = rng.normal(size=n) * 1e3
    s = ["%d %d" % (i, i + 1) if i % 7 else "" for i in range(n)]
    p = str(tmp_path / "t.dbf")
    shapefile.write_dbf(p, [("A", "N", 9, 0), ("X", "N", 24, 15), ("S", "C", 20, 0)], {"A": a, "X": x, "S": s})
    t = shapefile.DbfTable(p, update=True)
    assert np.array_equal(t.column_int("A"), a)
    np.testing.assert_allclose(t.column_float("X"), x, rtol=0, atol=1e-12)           # 15 decimals
    assert t.column_str("S") == s
    assert all(t.value(i, "A") == int(a[i]) for i in range(0, n, 613))
    y = rng.normal(size=n)
    t.set_column("X", y)
    t.set_column("A", np.arange(3), rows=[5, 6, 7])
    t.flush()
    u = shapefile.DbfTable(p)
    np.testing.assert_allclose(u.column_float("X"), y, rtol=0, atol=1e-15)
    assert u.column_int("A")[5:8].tolist() == [0, 1, 2] and u.column_int("A")[8] == a[8]
    with pytest.raises(ValueError):
        t.set_column("A", np.array([10 ** 12]), rows=[0])                           # does not fit N(9,0)
    with pytest.raises(ValueError, match="Can not open"):
        shapefile.DbfTable(str(tmp_path / "missing.dbf"))
    assert shapefile.Open(str(tmp_path / "missing.shp")) is None                      # OGR returns None


def test_point_geometry(tmp_path):
    # hand-built point .shp: 100-byte header (file code 9994 big endian, shape type little endian), then
    # records of [number BE, content length BE in 16-bit words][type LE, X, Y]
    pts = [(500000.5, 4100000.25), (500010.0, 4099990.0), (-3.5, 7.0)]
    body = b"".join(struct.pack(">ii", i + 1, 10) + struct.pack("<idd", 1, x, y) for i, (x, y) in enumerate(pts))
    head = bytearray(100)
    struct.pack_into(">i", head, 0, 9994)
    struct.pack_into(">i", head, 24, (100 + len(body)) // 2)
    struct.pack_into("<ii", head, 28, 1000, 1)
    (tmp_path / "PointsGCS.shp").write_bytes(bytes(head) + body)
    shapefile.write_dbf(str(tmp_path / "PointsGCS.dbf"), [("inner", "N", 9, 0), ("object", "N", 9, 0)],
                        {"inner": [32, 32, 16], "object": [64, 80, 64]})
    X, Y = shapefile.read_points(str(tmp_path / "PointsGCS.shp"))
    assert X.tolist() == [p[0] for p in pts] and Y.tolist() == [p[1] for p in pts]
    layer = shapefile.Open(str(tmp_path / "PointsGCS.shp")).GetLayer(0)
    f = layer.GetFeature(1)
    g = f.GetGeometryRef()
    assert (g.GetX(), g.GetY()) == pts[1] and int(f.GetField("object")) == 80
    # the reference's pixel mapping on these coordinates (MyUtils2.py:241-242), vectorised
    gt = (499990.0, 0.5, 0.0, 4100010.0, 0.0, -0.5)
    xp, yl = MyUtils2.geo_to_pixel(gt, X[:2], Y[:2])
    assert xp.tolist() == [int(abs((gt[0] - x) / gt[1]) + 1) for x, _ in pts[:2]]
    assert yl.tolist() == [int(abs((gt[3] - y) / gt[5]) + 1) for _, y in pts[:2]]
    # the module's own writer produces the same coordinates back
    shapefile.write_point_shp(str(tmp_path / "w.shp"), X, Y)
    X2, Y2 = shapefile.read_points(str(tmp_path / "w.shp"))
    assert np.array_equal(X, X2) and np.array_equal(Y, Y2)
    with pytest.raises(ValueError, match="Can not open"):
        shapefile.read_points(str(tmp_path / "PointsGCS.dbf"))


def test_create_field_and_set_feature_like_the_reference(tmp_path):
    """ExtractFeatures.py:181-186, 217-219: add OFTReal `simi` when missing, then GetFeature / SetField / SetFeature."""
    p = str(tmp_path / "lines.dbf")
    shapefile.write_dbf(p, [("LEFT_FID", "N", 9, 0), ("RIGHT_FID", "N", 9, 0)], {"LEFT_FID": [0, 1, 2], "RIGHT_FID": [1, 2, 0]})
    ro = shapefile.ShapefileLayer(str(tmp_path / "lines.shp"), 0)
    assert ro.CreateField(shapefile.FieldDefn("simi", shapefile.OFTReal), 1) != 0        # read-only: refused
    layer = shapefile.Open(str(tmp_path / "lines.shp"), 1).GetLayer(0)
    assert layer.GetLayerDefn().GetFieldIndex("simi") < 0
    assert layer.CreateField(shapefile.FieldDefn("simi", shapefile.OFTReal), 1) == 0
    d = layer.GetLayerDefn().GetFieldDefn(layer.GetLayerDefn().GetFieldIndex("simi"))
    assert (d.type, d.width, d.decimals) == ("N", 24, 15)                               # what the OGR driver creates
    f = layer.GetFeature(1)
    assert f.GetField("simi") is None
    f.SetField("simi", float(np.float32(0.7310586)))
    assert layer.SetFeature(f) == 0
    again = shapefile.ShapefileLayer(str(tmp_path / "lines.shp"))
    assert again.GetFeature(1).GetField("simi") == pytest.approx(float(np.float32(0.7310586)), abs=1e-15)
    assert again.GetFeature(0).GetField("simi") is None and again.GetFeature(1).GetField("LEFT_FID") == 1
    assert os.path.getsize(p) == 32 + 32 * 3 + 1 + 3 * (1 + 9 + 9 + 24) + 1


def test_edge_reader_on_real_files_matches_reference_golden(golden_dir, tmp_path, capsys):
    """The reference's PolygonConnectPointDataset rows (golden, produced by executing it) from actual
    shapefile attribute tables on disk, opened by path without GDAL."""
    e = np.load(os.path.join(golden_dir, "edge_reader.npz"))
    d = tmp_path / "tileA"
    d.mkdir()
    shapefile.write_dbf(str(d / "lines.dbf"), [("LEFT_FID", "N", 9, 0), ("RIGHT_FID", "N", 9, 0)],
                        {"LEFT_FID": e["left"], "RIGHT_FID": e["right"]})
    shapefile.write_dbf(str(tmp_path / "tileA.dbf"), [("PointID", "C", 40, 0)], {"PointID": ["1 2", "3"]})
    shapefile.write_dbf(str(d / "PointsGCS.dbf"), [("inner", "N", 9, 0)], {"inner": [1, 2, 3, 4]})
    ds = MyUtils2.PolygonConnectPointDataset(None, str(tmp_path / "tileA.shp"), str(d / "lines.shp"), str(d / "PointsGCS.shp"))
    rows = [ds[i] for i in range(len(ds))]
    assert [r[0] for r in rows] == e["out_fid"].tolist()
    assert [r[2] for r in rows] == e["out_left"].tolist() and [r[3] for r in rows] == e["out_right"].tolist()
    assert isinstance(ds.line_layer, shapefile.ShapefileLayer) and ds.polygon_layer.GetFeature(0).GetField("PointID") == "1 2"


def cpu_scorer(store, fields, left, right):
    """np.mean + the reference's expanded distance formula on the host (a stand-in for the GPU scorer)."""
    off, ids = membership_csr(fields)
    means = np.stack([store[ids[off[r]:off[r + 1]]].mean(axis=0) if off[r + 1] > off[r] else np.zeros(store.shape[1], np.float32)
                      for r in range(len(fields))]).astype(np.float32)
    x, y = means[np.asarray(left)], means[np.asarray(right)]
    d2 = (x * x).sum(1) + (y * y).sum(1) - 2 * (x * y).sum(1)
    return means, np.sqrt(np.maximum(d2, 0)).astype(np.float64)


class OgrOnly:
    """Hides the adaptor type so that score_layers takes its generic feature-by-feature OGR path."""

    def __init__(self, layer):
        self._l = layer

    def __getattr__(self, k):
        return getattr(self._l, k)


@pytest.mark.parametrize("generic", [False, True])
def test_score_layers_writes_simi_for_every_kept_line(golden_dir, tmp_path, generic, monkeypatch):
    # the file handling is what is under test here: the GPU scorer is replaced by a host stand-in (the real one
    # runs in tests/test_gpu_mirrors.py::test_test_for_shp_on_real_shapefiles)
    import deepmerge_b200.ExtractFeatures as E
    monkeypatch.setattr(E, "pool_and_score", cpu_scorer)
    p = np.load(os.path.join(golden_dir, "pool_score.npz"))
    fields, left, right = list(p["fields"]), p["left"].copy(), p["right"].copy()
    left[3] = -1                                                                      # a line on the tile border: skipped
    shapefile.write_dbf(str(tmp_path / "poly.dbf"), [("PointID", "C", 60, 0)], {"PointID": fields})
    shapefile.write_dbf(str(tmp_path / "lines.dbf"), [("LEFT_FID", "N", 9, 0), ("RIGHT_FID", "N", 9, 0)],
                        {"LEFT_FID": left, "RIGHT_FID": right})
    poly = shapefile.ShapefileLayer(str(tmp_path / "poly.shp"))
    lines = shapefile.ShapefileLayer(str(tmp_path / "lines.shp"), 1)
    fids, l, r, simi = score_layers(p["store"], OgrOnly(poly) if generic else poly, OgrOnly(lines) if generic else lines)
    keep = np.arange(len(left)) != 3
    assert fids.tolist() == np.nonzero(keep)[0].tolist() and np.array_equal(l, left[keep]) and np.array_equal(r, right[keep])
    np.testing.assert_allclose(simi, p["simi"][keep], rtol=1e-3)                      # the executed reference's scores
    t = shapefile.DbfTable(str(tmp_path / "lines.dbf"))
    col = t.column_float("simi")
    assert np.isnan(col[3])
    np.testing.assert_allclose(col[keep], simi, rtol=0, atol=1e-15)


def test_feature_store_npy(tmp_path):
    store = np.random.default_rng(0).normal(size=(17, 100)).astype(np.float32)
    np.save(tmp_path / "feats.npy", store)
    io = FeatureIO()
    io.ReadFeatures(str(tmp_path / "feats.npy"))
    assert np.array_equal(io.GetFeaturesByID(5), store[5]) and io.GetFeaturesByID(5).shape == (100,)
    with pytest.raises(IndexError):
        io.GetFeaturesByID(17)
    io.Close()


def rect_partition(H, W, rng, min_side=3):
    """Random partition of an H x W raster into axis-aligned rectangles -> (labels, [(r0, r1, c0, c1), ...])."""
    rects, todo = [], [(0, H, 0, W)]
    while todo:
        r0, r1, c0, c1 = todo.pop()
        h, w = r1 - r0, c1 - c0
        if (h < 2 * min_side and w < 2 * min_side) or rng.random() < 0.15:
            rects.append((r0, r1, c0, c1))
        elif w >= h and w >= 2 * min_side:
            m = int(rng.integers(c0 + min_side, c1 - min_side + 1))
            todo += [(r0, r1, c0, m), (r0, r1, m, c1)]
        elif h >= 2 * min_side:
            m = int(rng.integers(r0 + min_side, r1 - min_side + 1))
            todo += [(r0, m, c0, c1), (m, r1, c0, c1)]
        else:
            rects.append((r0, r1, c0, c1))
    labels = np.zeros((H, W), np.int32)
    for i, (r0, r1, c0, c1) in enumerate(rects):
        labels[r0:r1, c0:c1] = i
    return labels, rects


def test_polygon_shapefile_to_label_raster(tmp_path):
    """Polygons on disk -> the int32 label raster the raster-native path starts from (label = FID)."""
    rng = np.random.default_rng(5)
    H, W = 90, 130
    gt = (4000.0, 2.0, 0.0, 7000.0, 0.0, -2.0)
    labels, rects = rect_partition(H, W, rng)
    polys = [[np.array([[gt[0] + c0 * gt[1], gt[3] + r0 * gt[5]], [gt[0] + c1 * gt[1], gt[3] + r0 * gt[5]],
                        [gt[0] + c1 * gt[1], gt[3] + r1 * gt[5]], [gt[0] + c0 * gt[1], gt[3] + r1 * gt[5]]])]
             for r0, r1, c0, c1 in rects]
    shapefile.write_polygon_shp(str(tmp_path / "tile.shp"), polys)
    shapefile.write_dbf(str(tmp_path / "tile.dbf"), [("PointID", "C", 20, 0)], {"PointID": ["%d" % i for i in range(len(polys))]})
    back = shapefile.read_polygons(str(tmp_path / "tile.shp"))
    assert len(back) == len(polys) and all(np.array_equal(b[0][:4], p[0]) for b, p in zip(back, polys))
    got = shapefile.rasterize_polygons(back, gt, H, W)
    assert got.dtype == np.int32 and np.array_equal(got, labels)          # pixel-aligned cells come back exactly
    # a window of the same scene (shifted origin), ids from an attribute instead of the FID
    sub = shapefile.rasterize_polygons(back, (gt[0] + 10 * gt[1], gt[1], 0.0, gt[3] + 20 * gt[5], 0.0, gt[5]), 30, 40,
                                       ids=np.arange(len(polys)) + 1000)
    assert np.array_equal(sub, labels[20:50, 10:50] + 1000)
    with pytest.raises(ValueError):
        shapefile.rasterize_polygons(back, (0, 1, 0.1, 0, 0, -1), 4, 4)
    with pytest.raises(ValueError, match="not a polygon"):
        shapefile.read_polygons(_point_file(tmp_path))


def _point_file(tmp_path):
    shapefile.write_point_shp(str(tmp_path / "pts.shp"), [1.0], [2.0])
    return str(tmp_path / "pts.shp")


def test_rasterize_holes_overlaps_and_brute_force(tmp_path):
    # a ring inside a ring is a hole (even-odd); the later polygon wins where two overlap; uncovered pixels stay nodata
    sq = lambda a, b: np.array([[a, a], [b, a], [b, b], [a, b]], float)
    gt = (0.0, 1.0, 0.0, 12.0, 0.0, -1.0)
    lab = shapefile.rasterize_polygons([[sq(1, 11), sq(4, 8)], [sq(5, 7)], []], gt, 12, 12)
    want = np.full((12, 12), -1, np.int32)
    want[1:11, 1:11] = 0
    want[4:8, 4:8] = -1
    want[5:7, 5:7] = 1
    assert np.array_equal(lab, want)
    # random (non-convex) polygons against a per-pixel even-odd test
    rng = np.random.default_rng(1)
    H, W = 50, 70
    polys = []
    for _ in range(25):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        ang = np.sort(rng.uniform(0, 2 * np.pi, int(rng.integers(3, 10))))
        rad = rng.uniform(1, 12, len(ang))
        polys.append([np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)])
    got = shapefile.rasterize_polygons(polys, (0.0, 1.0, 0.0, 0.0, 0.0, 1.0), H, W)
    cc, rr = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    want = np.full((H, W), -1, np.int32)
    for fid, (ring,) in enumerate(polys):
        inside = np.zeros((H, W), bool)
        for (a, b), (c, d) in zip(ring, np.roll(ring, -1, axis=0)):
            cross = (b <= rr) != (d <= rr)
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = a + (rr - b) * (c - a) / (d - b)
            inside ^= cross & (cc >= xi)
        want[inside] = fid
    assert np.array_equal(got, want)
    # PolygonZ records (type 15) carry the same header; round trip through the writer's type-5 layout patched to 15
    shapefile.write_polygon_shp(str(tmp_path / "z.shp"), polys[:3])
    raw = bytearray((tmp_path / "z.shp").read_bytes())
    struct.pack_into("<i", raw, 32, 15)
    off = 100
    while off < len(raw):
        (clen,) = struct.unpack_from(">i", raw, off + 4)
        struct.pack_into("<i", raw, off + 8, 15)
        off += 8 + 2 * clen
    (tmp_path / "z.shp").write_bytes(bytes(raw))
    z = shapefile.read_polygons(str(tmp_path / "z.shp"))
    assert all(np.allclose(a[0][:-1], b[0]) for a, b in zip(z, polys[:3]))


def test_region_of_point_from_point_id_fields():
    from deepmerge_b200.ExtractFeatures import region_of_point
    rop = region_of_point(["3 1", "", "0 4"], 6)
    assert rop.dtype == np.int32 and rop.tolist() == [2, 0, -1, 0, 2, -1]
    with pytest.raises(ValueError):
        region_of_point(["7"], 3)
