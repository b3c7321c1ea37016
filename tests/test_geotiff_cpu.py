"""CPU: the GDAL-free raster adaptor against Pillow as an independent TIFF
encoder / decoder, hand-built BigTIFF bytes, and the reference's window conventions."""
import struct

import numpy as np
import pytest

from deepmerge_b200 import MyUtils2, geotiff

PIL = pytest.importorskip("PIL")
from PIL import Image, TiffImagePlugin      # noqa: E402


def scene(H, W, C, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (H, W, C), dtype=np.uint8)


@pytest.mark.parametrize("C,tile,deflate", [(3, None, False), (4, None, True), (1, 64, False), (4, 128, True)])
def test_own_writer_is_read_by_pillow_and_by_the_adaptor(tmp_path, C, tile, deflate):
    a = scene(301, 517, C, seed=C)
    gt = (500000.0, 0.5, 0.0, 4100000.0, 0.0, -0.5)
    p = str(tmp_path / "s.tif")
    geotiff.write_geotiff(p, a, gt, tile=tile, deflate=deflate)
    with Image.open(p) as im:                                   # independent decoder
        b = np.asarray(im)
    assert np.array_equal(b.reshape(a.shape), a)
    ds = geotiff.Open(p)
    assert (ds.RasterCount, ds.RasterYSize, ds.RasterXSize) == (C, 301, 517)
    assert ds.GetGeoTransform() == gt
    full = ds.ReadAsArray()
    assert np.array_equal(full if C > 1 else full[None], np.moveaxis(a, 2, 0))
    assert np.array_equal(ds.interleaved(), a)


@pytest.mark.parametrize("compression", [None, "tiff_adobe_deflate", "tiff_lzw"])
def test_pillow_written_geotiff(tmp_path, compression):
    a = scene(200, 333, 3, seed=7)
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550] = (2.0, 2.0, 0.0)
    ifd.tagtype[33550] = 12
    ifd[33922] = (0.0, 0.0, 0.0, 1000.0, 9000.0, 0.0)
    ifd.tagtype[33922] = 12
    p = str(tmp_path / "p.tif")
    Image.fromarray(a).save(p, tiffinfo=ifd, **({"compression": compression} if compression else {}))
    ds = geotiff.Open(p)
    assert ds is not None and np.array_equal(ds.interleaved(), a)
    assert ds.GetGeoTransform() == (1000.0, 2.0, 0.0, 9000.0, 0.0, -2.0)
    # window reads: GDAL layout [C, h, w]; out-of-range requests fail
    w = ds.ReadAsArray(10, 20, 30, 40)
    assert w.shape == (3, 40, 30) and np.array_equal(w, np.moveaxis(a[20:60, 10:40], 2, 0))
    assert ds.ReadAsArray(300, 0, 64, 64) is None


def test_bigtiff_pixel_is_point_world_file_and_npy(tmp_path):
    # 3 x 2 single-band BigTIFF laid out by hand: header (II, 43, 8, 0, first IFD offset), 20-byte entries
    px = bytes([1, 2, 3, 4, 5, 6])
    ents = [(256, 4, 1, 3), (257, 4, 1, 2), (258, 3, 1, 8), (259, 3, 1, 1), (262, 3, 1, 1), (273, 16, 1, 16),
            (277, 3, 1, 1), (278, 4, 1, 2), (279, 16, 1, 6)]
    raw = struct.pack("<2sHHHQ", b"II", 43, 8, 0, 24) + px + b"\0\0"
    raw += struct.pack("<Q", len(ents))
    for tag, typ, n, v in ents:
        raw += struct.pack("<HHQQ", tag, typ, n, v)
    raw += struct.pack("<Q", 0)
    p = tmp_path / "big.tif"
    p.write_bytes(raw)
    ds = geotiff.Open(str(p))
    assert ds.RasterCount == 1 and ds.ReadAsArray().tolist() == [[1, 2, 3], [4, 5, 6]]
    assert ds.GetGeoTransform() == (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)                     # GDAL's default
    # world file beside the image: centre of the upper-left pixel -> corner
    (tmp_path / "big.tfw").write_text("2.0\n0.0\n0.0\n-2.0\n101.0\n899.0\n")
    assert geotiff.Open(str(p)).GetGeoTransform() == (100.0, 2.0, 0.0, 900.0, 0.0, -2.0)
    # PixelIsPoint GeoTIFF: GDAL shifts the origin by half a pixel
    tags = {33550: (2.0, 2.0, 0.0), 33922: (0.0, 0.0, 0.0, 101.0, 899.0, 0.0), 34735: (1, 1, 0, 1, 1025, 0, 1, 2)}
    assert geotiff._geotransform_from_tags(tags) == (100.0, 2.0, 0.0, 900.0, 0.0, -2.0)
    tags[34735] = (1, 1, 0, 1, 1025, 0, 1, 1)
    assert geotiff._geotransform_from_tags(tags) == (101.0, 2.0, 0.0, 899.0, 0.0, -2.0)
    # .npy scenes and failures
    a = scene(20, 30, 4)
    np.save(tmp_path / "s.npy", a)
    assert np.array_equal(geotiff.Open(str(tmp_path / "s.npy")).interleaved(), a)
    assert geotiff.Open(str(tmp_path / "missing.tif")) is None
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    assert geotiff.Open(str(tmp_path / "junk.tif")) is None


def test_reference_window_cut_on_an_opened_raster(tmp_path):
    """cut_image (MyUtils2.py:330-360) and the sample-point pixel mapping (:241-242) on a file opened without GDAL."""
    a = scene(120, 150, 3, seed=3)
    gt = (1000.0, 0.5, 0.0, 5000.0, 0.0, -0.5)
    geotiff.write_geotiff(str(tmp_path / "img.tif"), a, gt)
    ds = geotiff.Open(str(tmp_path / "img.tif"))
    xp, yl = MyUtils2.geo_to_pixel(ds.GetGeoTransform(), np.array([1000.2, 1070.0]), np.array([4999.9, 4945.0]))
    assert xp.tolist() == [1, 141] and yl.tolist() == [1, 111]
    for x, y in zip(xp.tolist(), yl.tolist()):
        win = MyUtils2.calculate_left_top_point_and_size(x, y, 64)
        got = MyUtils2.cut_image(ds, win)
        want = np.zeros((3, 64, 64), np.uint8)
        x0, y0 = win[0], win[1]
        xs, ys = max(x0, 0), max(y0, 0)
        xe, ye = min(x0 + 64, 150), min(y0 + 64, 120)
        want[:, ys - y0:ye - y0, xs - x0:xe - x0] = np.moveaxis(a[ys:ye, xs:xe], 2, 0)
        assert np.array_equal(got, want)
        assert np.array_equal(got, MyUtils2.cut_image(ds.array, win))                  # array and dataset forms agree


def test_extract_feature_dataset_on_real_files(tmp_path):
    """ExtractFeatureDataset (MyUtils2.py:213-437) from a GeoTIFF + PointsGCS.shp on disk, without GDAL: designed
    attributes + scale factors, scales, pixel positions and the per-scale zero-padded cuts, against the reference's
    per-feature formulas restated inline."""
    from deepmerge_b200 import shapefile
    rng = np.random.default_rng(11)
    a = scene(160, 200, 3, seed=5)
    gt = (3000.0, 2.0, 0.0, 8000.0, 0.0, -2.0)
    geotiff.write_geotiff(str(tmp_path / "img.tif"), a, gt, tile=64, deflate=True)
    n = 9
    X = gt[0] + rng.uniform(0, 200 * 2.0, n)
    Y = gt[3] - rng.uniform(0, 160 * 2.0, n)
    shapefile.write_point_shp(str(tmp_path / "PointsGCS.shp"), X, Y)
    cols = {f: np.round(rng.uniform(0, 500, n), 6) for f in MyUtils2.DESIGNED_FIELDS}
    cols["inner"], cols["object"] = rng.integers(8, 40, n), rng.integers(40, 90, n)
    fields = [(f, "N", 24, 15) for f in MyUtils2.DESIGNED_FIELDS] + [("inner", "N", 9, 0), ("object", "N", 9, 0)]
    shapefile.write_dbf(str(tmp_path / "PointsGCS.dbf"), fields, cols)
    ds = MyUtils2.ExtractFeatureDataset(str(tmp_path / "img.tif"), str(tmp_path / "PointsGCS.shp"))
    assert len(ds) == n and ds.band_num == 3
    w = ds.windows()
    for k in range(n):
        inner, obj = int(cols["inner"][k]), int(cols["object"][k])
        scales, factors = MyUtils2.get_scales(inner, obj)                # pinned against the executed reference (golden)
        assert w["scales"][k].tolist() == scales
        want = np.asarray([float(cols[f][k]) for f in MyUtils2.DESIGNED_FIELDS] + factors, np.float32)
        np.testing.assert_array_equal(w["designed"][k], want)
        xp = int(abs((gt[0] - X[k]) / gt[1]) + 1)
        yl = int(abs((gt[3] - Y[k]) / gt[5]) + 1)
        assert (int(w["xpix"][k]), int(w["ylin"][k])) == (xp, yl)
        patches = ds.window_patches(k)
        assert [p.shape for p in patches] == [(3, s, s) for s in scales]
        for p, s in zip(patches, scales):
            assert np.array_equal(p, MyUtils2.cut_image(np.moveaxis(a, 2, 0), MyUtils2.calculate_left_top_point_and_size(xp, yl, s)))
    # the generic feature-by-feature path (what runs on real OGR layers) gives the same arrays
    class OgrOnly:
        def __init__(self, layer):
            self._l = layer

        def ResetReading(self):
            self._l.ResetReading()

        def GetNextFeature(self):
            return self._l.GetNextFeature()

        def GetFeature(self, i):
            return self._l.GetFeature(i)

    generic = MyUtils2.ExtractFeatureDataset(None, None, img_dataset=ds.img_dataset, point_layer=OgrOnly(ds.point_layers))
    g = generic.windows()
    assert all(np.array_equal(g[k], w[k]) for k in w)
    with pytest.raises(ValueError, match="Can not open"):
        MyUtils2.ExtractFeatureDataset(str(tmp_path / "nope.tif"), str(tmp_path / "PointsGCS.shp"))
