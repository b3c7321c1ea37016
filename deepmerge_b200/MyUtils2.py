"""Drop-in for the RAG edge-list reader of the reference's MyUtils2.py.

PolygonConnectPointDataset (MyUtils2.py:128-209) iterates the polyline layer of lines.shp and
keeps [line_fid, tile_name, LEFT_FID, RIGHT_FID] for every feature whose two ids are not -1
(:184-186).  With osgeo importable it opens the paths exactly like the reference; without it
(this image) the shapefiles are opened by deepmerge_b200.shapefile, a pure-Python reader of the
.dbf / .shp parts this path touches.  Already-opened layer objects (anything with ResetReading /
GetNextFeature / GetField / GetFID) can be passed instead of paths.  `edge_keys()` hands the list
to the GPU path as packed (min,max) keys.
"""
from __future__ import annotations

import numpy as np


class PolygonConnectPointDataset:
    def __init__(self, image_path, polygon_path, polyline_path, point_path, num=0, *, polygon_layer=None,
                 line_layer=None, point_layer=None, img_dataset=None):
        self.image_path, self.polygon_path = image_path, polygon_path
        self.polyline_path, self.point_path, self.num = polyline_path, point_path, num
        self.data = []
        self.polygon_dataset = self.point_dataset = self.line_dataset = None
        self.img_dataset = img_dataset
        self.polygon_layer, self.point_layer, self.line_layer = polygon_layer, point_layer, line_layer
        if line_layer is None:
            self._open_with_ogr()
        count = self.add_data(self.polygon_path)
        print("OK ", count)

    def _open_with_ogr(self):
        try:
            from osgeo import gdal, ogr
        except ImportError:
            return self._open_without_gdal()
        drv = ogr.GetDriverByName("ESRI Shapefile")
        for attr, path in (("polygon", self.polygon_path), ("point", self.point_path), ("line", self.polyline_path)):
            ds = drv.Open(path, 1)
            if ds is None or ds.GetLayer(0) is None:
                raise ValueError("Can not open {0}".format(path))
            setattr(self, attr + "_dataset", ds)
            setattr(self, attr + "_layer", ds.GetLayer(0))
        self.img_dataset = gdal.Open(self.image_path, gdal.GA_ReadOnly)
        if self.img_dataset is None:
            raise ValueError("Can not open {0}".format(self.image_path))

    def _open_without_gdal(self):
        """osgeo is not installed: the attribute tables (and point coordinates) are read by the pure-Python
        adaptor deepmerge_b200.shapefile, which offers the OGR calls this path makes, and the image (only used
        for its band count here) by deepmerge_b200.geotiff."""
        from . import shapefile
        for attr, path in (("polygon", self.polygon_path), ("point", self.point_path), ("line", self.polyline_path)):
            ds = shapefile.Open(path, 1) if path else None
            if ds is None:
                if attr == "line" or path:
                    raise ValueError("Can not open {0}".format(path))
                continue
            setattr(self, attr + "_dataset", ds)
            setattr(self, attr + "_layer", ds.GetLayer(0))
        if self.img_dataset is None and self.image_path:
            from . import geotiff
            self.img_dataset = geotiff.Open(self.image_path)
            if self.img_dataset is None:
                raise ValueError("Can not open {0}".format(self.image_path))

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.data[index]

    def add_data(self, txt_path):
        if txt_path is None:
            return None
        name = txt_path.split("\\")[-1].split(".")[0]
        if self.img_dataset is not None:
            self.band_num = self.img_dataset.RasterCount
        count = 0
        self.line_layer.ResetReading()
        feature = self.line_layer.GetNextFeature()
        while feature is not None:
            left, right = int(feature.GetField("LEFT_FID")), int(feature.GetField("RIGHT_FID"))
            if left != -1 and right != -1:
                self.data.append([int(feature.GetFID()), name, left, right])
                count += 1
            feature = self.line_layer.GetNextFeature()
        return count

    def edge_keys(self):
        """Packed (min << 32) | max keys of the kept rows, in file order (duplicates kept)."""
        lr = np.asarray([[d[2], d[3]] for d in self.data], np.int64).reshape(-1, 2)
        return (np.minimum(lr[:, 0], lr[:, 1]) << 32) | np.maximum(lr[:, 0], lr[:, 1])


# --------------------------------------------------------------------------------------------
# Geometry conventions of ExtractFeatureDataset (MyUtils2.py:213-437) that the raster path must
# reproduce when it maps sample points to pixels.  Vectorised over
# points; integer truncation follows Python's int() (toward zero) like the reference.
# --------------------------------------------------------------------------------------------
SCALES = (32, 64, 128, 1)          # config.py:32 `configs.scales`


def geo_to_pixel(geo_transform, X, Y):
    """Sample-point coordinates -> (XPixel, YLine) as the reference computes them, MyUtils2.py:241-242:
    int(abs((gt[0] - X) / gt[1]) + 1) -- note the +1 -- and the same with gt[3], gt[5] for rows."""
    gt = geo_transform
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    px = np.trunc(np.abs((gt[0] - X) / gt[1]) + 1).astype(np.int64)
    ln = np.trunc(np.abs((gt[3] - Y) / gt[5]) + 1).astype(np.int64)
    return px, ln


def calculate_left_top_point_and_size(midPointX, midPointY, windowLength):
    """Window (left, top, w, h) centred on a pixel, MyUtils2.py:379-383: int(mid - w / 2) truncates toward zero."""
    mx, my, w = (np.asarray(v) for v in (midPointX, midPointY, windowLength))
    left = np.trunc(mx - w / 2).astype(np.int64)
    top = np.trunc(my - w / 2).astype(np.int64)
    w = np.trunc(w).astype(np.int64)
    if left.ndim == 0:
        return int(left), int(top), int(w), int(w)
    return left, top, w, w


def get_scales(inner_scale, object_scale, cfg_scales=SCALES):
    """The four window sizes [inner, object, object + d, object + 2 d], d = int(object - inner), and their
    factors size / configs.scales[i] (MyUtils2.py:300-327)."""
    interval = int(object_scale - inner_scale)
    scales = [inner_scale, object_scale, object_scale + interval, object_scale + 2 * interval]
    factors = [s * 1.0 / c for s, c in zip(scales, cfg_scales)]
    return scales, factors


def cut_image(image, window):
    """Zero-padded window of a [C, H, W] uint8 raster, MyUtils2.py:330-360, including the reference's clamp
    `start + size >= src -> size = src - start` (so a window that ends exactly at the border is cut the same).
    `image` is a numpy array or a GDAL-like dataset (RasterCount / RasterXSize / RasterYSize / ReadAsArray), as in
    the reference -- e.g. what deepmerge_b200.geotiff.Open returns."""
    x0, y0, w, h = (int(v) for v in window)
    dataset = image if hasattr(image, "ReadAsArray") else None
    C, H, W = (image.RasterCount, image.RasterYSize, image.RasterXSize) if dataset is not None else image.shape
    out = np.zeros((C, h, w), np.uint8)
    ox = oy = 0
    dw, dh = w, h
    if x0 < 0:
        dw += x0
        ox, x0 = -x0, 0
    if x0 + dw >= W:
        dw = W - x0
    if y0 < 0:
        dh += y0
        oy, y0 = -y0, 0
    if y0 + dh >= H:
        dh = H - y0
    if dw > 0 and dh > 0:
        out[:, oy:oy + dh, ox:ox + dw] = (dataset.ReadAsArray(x0, y0, dw, dh) if dataset is not None
                                          else image[:, y0:y0 + dh, x0:x0 + dw])
    return out


DESIGNED_FIELDS = ("area", "peri", "len", "width", "smooth", "std0", "std1", "std2", "mean0", "mean1", "mean2",
                   "shapeness", "compact", "bright", "border")       # get_designed_features, MyUtils2.py:250-284


class ExtractFeatureDataset:
    """The sample-point side of the reference's ExtractFeatureDataset (MyUtils2.py:213-437): one item per feature
    of PointsGCS.shp -- its 15 designed attributes, the four window scales from `inner` / `object`, and its pixel
    position in the scene raster.  Files are opened with GDAL/OGR when installed, else with the adaptors
    deepmerge_b200.shapefile / deepmerge_b200.geotiff.  `windows()` returns everything for all points as arrays (what
    the batched GPU cutter `cut_windows` takes); `window_patches(i)` is the per-point, per-scale zero-padded cut of
    get_patches_by_scales BEFORE its INTER_AREA resize (the resize and the network are the missing part of N1)."""

    def __init__(self, image_path, point_path, *, img_dataset=None, point_layer=None):
        self.image_path, self.point_path = image_path, point_path
        self.data = []
        self._windows = None
        self.img_dataset, self.point_layers, self.point_dataset = img_dataset, point_layer, None
        self.add_data(image_path, point_path)

    def __len__(self):
        return len(self.data)

    def add_data(self, img_path, point_path):
        if self.point_layers is None or self.img_dataset is None:
            if point_path is None or img_path is None:
                return None
            try:
                from osgeo import gdal, ogr
                ds = ogr.GetDriverByName("ESRI Shapefile").Open(point_path, 1)
                img = gdal.Open(img_path, gdal.GA_ReadOnly)
            except ImportError:
                from . import geotiff, shapefile
                ds, img = shapefile.Open(point_path, 1), geotiff.Open(img_path)
            if ds is None or ds.GetLayer(0) is None:
                raise ValueError("Can not open {0}".format(point_path))
            if img is None:
                raise ValueError("Can not open {0}".format(img_path))
            self.point_dataset, self.point_layers, self.img_dataset = ds, ds.GetLayer(0), img
        self.band_num = self.img_dataset.RasterCount
        self.point_layers.ResetReading()
        feature = self.point_layers.GetNextFeature()
        while feature is not None:
            self.data.append(int(feature.GetFID()))
            feature = self.point_layers.GetNextFeature()
        return len(self.data)

    def windows(self):
        """-> dict of arrays over all points: ids int64 [N], designed float32 [N, 19] (15 attributes + 4 scale
        factors, as get_all_features concatenates them), scales int64 [N, 4], xpix / ylin int64 [N]."""
        if self._windows is not None:
            return self._windows
        ids = np.asarray(self.data, np.int64)
        table = getattr(self.point_layers, "table", None)
        if table is not None and getattr(self.point_layers, "points", None) is not None:       # whole columns at once
            attr = np.stack([table.column_float(f)[ids] for f in DESIGNED_FIELDS], axis=1)
            inner, obj = table.column_int("inner")[ids], table.column_int("object")[ids]
            X, Y = (c[ids] for c in self.point_layers.points)
        else:
            feats = [self.point_layers.GetFeature(int(i)) for i in ids]
            attr = np.asarray([[float(f.GetField(n)) for n in DESIGNED_FIELDS] for f in feats], np.float64).reshape(-1, 15)
            inner = np.asarray([int(f.GetField("inner")) for f in feats], np.int64)
            obj = np.asarray([int(f.GetField("object")) for f in feats], np.int64)
            X = np.asarray([f.GetGeometryRef().GetX() for f in feats], np.float64)
            Y = np.asarray([f.GetGeometryRef().GetY() for f in feats], np.float64)
        interval = obj - inner                                           # get_scales, MyUtils2.py:300-327
        scales = np.stack([inner, obj, obj + interval, obj + 2 * interval], axis=1).astype(np.int64)
        factors = scales / np.asarray(SCALES, np.float64)
        xpix, ylin = geo_to_pixel(self.img_dataset.GetGeoTransform(), X, Y)
        self._windows = {"ids": ids, "designed": np.concatenate([attr, factors], axis=1).astype(np.float32),
                         "scales": scales, "xpix": xpix, "ylin": ylin}
        return self._windows

    def window_patches(self, index):
        """[uint8 [C, s, s] for s in the point's four scales]: cut_image at calculate_left_top_point_and_size."""
        w = self.windows()
        k = int(np.nonzero(w["ids"] == self.data[index])[0][0])
        return [cut_image(self.img_dataset, calculate_left_top_point_and_size(int(w["xpix"][k]), int(w["ylin"][k]), int(s)))
                for s in w["scales"][k]]


def cut_windows(image, xpix, ylin, size):
    """Batched cut_image on the GPU (dm_cut_windows): image uint8 CUDA tensor [C, H, W] (band-major, as GDAL reads
    it), pixel centres (xpix, ylin) -> uint8 [n, C, size, size] windows whose top-left corners follow
    calculate_left_top_point_and_size, zero-padded outside the raster (MyUtils2.py:330-383)."""
    import torch
    from ._lib import lib
    from .raster import _p, _stream
    if not image.is_cuda or image.dtype != torch.uint8 or image.dim() != 3:
        raise ValueError("image must be a uint8 CUDA tensor [C, H, W]")
    image = image.contiguous()
    C, H, W = image.shape
    left, top, _, _ = calculate_left_top_point_and_size(np.asarray(xpix), np.asarray(ylin), np.full(len(xpix), size))
    x0 = torch.as_tensor(left, dtype=torch.int32, device=image.device)
    y0 = torch.as_tensor(top, dtype=torch.int32, device=image.device)
    out = torch.empty((len(xpix), C, size, size), dtype=torch.uint8, device=image.device)
    L = lib()
    with torch.cuda.device(image.device):
        L.check(L.dm_cut_windows(_p(image), C, H, W, _p(x0), _p(y0), len(xpix), size, _p(out), _stream()), "dm_cut_windows")
    return out


def area_tables(s, t):
    """Coefficient tables of cv2.resize(.., (t, t), INTER_AREA) for a square source of side s (the layout
    dm_resize_area takes): -> (mode, ti int32 or None, tf float32 or None).  OpenCV's published rules: an integer
    factor needs no table (mode 0); a fractional shrink weighs every source cell by its overlap with the
    destination cell, float32 weights (mode 1); an enlargement is its "area-mode" bilinear with 11-bit
    fixed-point coefficients (mode 2)."""
    import math
    s, t = int(s), int(t)
    if s % t == 0:
        return 0, None, None
    f32 = np.float32
    if s > t:
        scale = s / t
        start, src, w = [0], [], []
        for d in range(t):
            lo = d * scale
            hi = lo + scale
            cell = min(scale, s - lo)
            i1 = math.ceil(lo)
            i2 = min(math.floor(hi), s - 1)
            i1 = min(i1, i2)
            if i1 - lo > 1e-3:
                src.append(i1 - 1)
                w.append(f32((i1 - lo) / cell))
            for i in range(i1, i2):
                src.append(i)
                w.append(f32(1.0 / cell))
            if hi - i2 > 1e-3:
                src.append(i2)
                w.append(f32(min(min(hi - i2, 1.0), cell) / cell))
            start.append(len(src))
        return 1, np.asarray(start + src, np.int32), np.asarray(w, np.float32)
    inv = t / s
    scale = 1.0 / inv
    sx, a0, a1, xmax = [], [], [], t
    for d in range(t):
        i = math.floor(d * scale)
        fx = f32((d + 1) - (i + 1) * inv)
        fx = f32(0) if fx <= 0 else f32(fx - math.floor(fx))
        if i + 1 >= s:
            xmax = min(xmax, d)
            if i >= s - 1:
                fx, i = f32(0), s - 1
        sx.append(i)
        a0.append(int(np.rint((f32(1) - fx) * f32(2048))))
        a1.append(int(np.rint(fx * f32(2048))))
    return 2, np.asarray(sx + a0 + a1 + [xmax], np.int32), None


def resize_windows(windows, t):
    """resize_data (MyUtils2.py:362-376) for a batch on the GPU: uint8 CUDA tensor [n, C, s, s] (e.g. from
    cut_windows) -> float32 [n, C, t, t] in [0, 1], bit-identical to cv2.resize(INTER_AREA) / 255 per band."""
    import torch
    from ._lib import lib
    from .raster import _p, _stream
    if not windows.is_cuda or windows.dtype != torch.uint8 or windows.dim() != 4 or windows.shape[2] != windows.shape[3]:
        raise ValueError("windows must be a uint8 CUDA tensor [n, C, s, s]")
    windows = windows.contiguous()
    n, C, s, _ = windows.shape
    mode, ti, tf = area_tables(s, t)
    ti_d = torch.from_numpy(ti).to(windows.device) if ti is not None else None
    tf_d = torch.from_numpy(tf).to(windows.device) if tf is not None else None
    out = torch.empty((n, C, int(t), int(t)), dtype=torch.float32, device=windows.device)
    L = lib()
    with torch.cuda.device(windows.device):
        L.check(L.dm_resize_area(_p(windows), n * C, s, int(t), mode, _p(ti_d), _p(tf_d), _p(out), _stream()), "dm_resize_area")
    return out


def point_patches(image, windows, cfg_scales=SCALES):
    """get_patches_by_scales (MyUtils2.py:286-298) for ALL points on the GPU: `image` uint8 CUDA tensor [C, H, W],
    `windows` = ExtractFeatureDataset.windows() -> list of 4 float32 tensors [N, C, cfg, cfg] (cfg = 32, 64, 128, 1).
    Points are grouped by their window size, each group is cut (cut_windows) and resized (resize_windows) at once."""
    import torch
    n = len(windows["ids"])
    xpix, ylin = np.asarray(windows["xpix"], np.int64), np.asarray(windows["ylin"], np.int64)
    out = []
    for k, cfg in enumerate(cfg_scales):
        sizes = np.asarray(windows["scales"])[:, k]
        patches = torch.empty((n, image.shape[0], cfg, cfg), dtype=torch.float32, device=image.device)
        for s in np.unique(sizes):
            sel = np.nonzero(sizes == s)[0]
            cut = cut_windows(image, xpix[sel], ylin[sel], int(s))
            patches[torch.as_tensor(sel, device=image.device)] = resize_windows(cut, cfg)
        out.append(patches)
    return out
