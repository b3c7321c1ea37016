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

import warnings

import numpy as np
import torch

from ._lib import lib
from .raster import _p, _stream, pool_points_csr, region_mean, score_l2


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("deepmerge_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def Euclidean_distance(X, Y):
    """X n*p, Y m*p -> D n*m with D[i,j] = sqrt(max(0, |X_i|^2 + |Y_j|^2 - 2 X_i.Y_j)).

    The arithmetic is float32 (what the reference computes for its float32 embeddings, ExtractFeatures.py:139-147).  The
    RESULT dtype follows the reference: float64 inputs give a float64 array -- holding values computed in float32, which
    is the one documented difference: the reference would carry float64 precision for float64 inputs."""
    out_dtype = np.result_type(np.asarray(X).dtype, np.asarray(Y).dtype, np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = np.ascontiguousarray(Y, dtype=np.float32)
    if X.ndim != 2 or Y.ndim != 2 or X.shape[1] != Y.shape[1]:
        raise ValueError("X and Y must be [n,p] and [m,p]")
    dev = _dev()
    L = lib()
    x, y = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    out = torch.empty((X.shape[0], Y.shape[0]), dtype=torch.float32, device=dev)
    L.check(L.dm_euclidean_matrix(_p(x), _p(y), X.shape[0], Y.shape[0], X.shape[1], _p(out), _stream()), "dm_euclidean_matrix")
    return out.cpu().numpy().astype(out_dtype, copy=False)


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


def region_of_point(point_id_fields, n_points, sep=" "):
    """'PointID' strings of polygons 0..R-1 -> int32 [n_points]: the polygon every sample point belongs to (-1 for a
    point no polygon lists) -- the membership form `merge_scene(region_of_point=...)` takes."""
    off, ids = membership_csr(point_id_fields, sep)
    rop = np.full(int(n_points), -1, np.int32)
    if ids.size and (ids.min() < 0 or ids.max() >= n_points):
        raise ValueError("PointID refers to a point outside the feature store")
    rop[ids] = np.repeat(np.arange(len(point_id_fields), dtype=np.int32), np.diff(off))
    return rop


def check_ids(point_ids, n_rows, left_ids, right_ids, n_polygons):
    """The kernels gather store rows by PointID and mean rows by polygon id without bounds tests: check on the host,
    and raise as the reference's GetFeaturesByID does for a row outside the store (ExtractFeatures.py:109-112)."""
    ids = np.asarray(point_ids).ravel()
    if ids.size and (int(ids.min()) < 0 or int(ids.max()) >= n_rows):
        bad = int(ids.max()) if int(ids.max()) >= n_rows else int(ids.min())
        raise IndexError("PointID %d is outside the feature store (%d rows)" % (bad, n_rows))
    lr = np.concatenate([np.asarray(left_ids, np.int64).ravel(), np.asarray(right_ids, np.int64).ravel()])
    if lr.size and (int(lr.min()) < 0 or int(lr.max()) >= n_polygons):
        bad = int(lr.max()) if int(lr.max()) >= n_polygons else int(lr.min())
        raise IndexError("LEFT_FID / RIGHT_FID %d is not a polygon id (%d polygons)" % (bad, n_polygons))


def pool_and_score(store, point_id_fields, left_ids, right_ids):
    """Mean-pool every polygon's member rows of the feature store (np.mean(axis=0) semantics, bit
    exact) and score every (left, right) edge with the Euclidean distance -> (means [R,D] float32,
    simi [E] float64, what the reference writes to the 'simi' OFTReal field :217-219)."""
    off, ids = membership_csr(point_id_fields)
    check_ids(ids, int(np.shape(store)[0]), left_ids, right_ids, len(point_id_fields))
    dev = _dev()
    store = np.ascontiguousarray(store, dtype=np.float32)
    with warnings.catch_warnings():                      # a read-only memory map is fine: the rows are only copied to the device
        warnings.simplefilter("ignore", UserWarning)
        store_t = torch.from_numpy(store).to(dev)
    s, c = pool_points_csr(torch.from_numpy(off).to(dev), torch.from_numpy(ids).to(dev), store_t)
    mean, n2 = region_mean(s, c)
    left = torch.as_tensor(np.asarray(left_ids, np.int64), device=dev)
    right = torch.as_tensor(np.asarray(right_ids, np.int64), device=dev)
    keys = (torch.minimum(left, right) << 32) | torch.maximum(left, right)
    simi = score_l2(mean, keys, n2)
    return mean.cpu().numpy(), simi.cpu().numpy().astype(np.float64)


class FeatureIO:
    """The feature-store half of the reference's FeatureIO (ExtractFeatures.py:88-116): the HDF5 dataset
    "dataset" [N, D] float32 that `extract_features` appends to and `test_for_shp` reads row by row.
    h5py is used when it is installed; a `.npy` file with the same [N, D] array is accepted as well
    (h5py is not part of this image).  The network half (`extract_features`) is out of scope here."""

    def __init__(self, net=None, checkpoint_path=None):
        self.net, self.checkpoint_path = net, checkpoint_path
        self.h5py_file = None
        self.dataset = None

    def save_h5(self, h5f, data, dataset_name="dataset"):
        """Append rows to the resizable dataset (created on first use), ExtractFeatures.py:88-101."""
        shape = list(data.shape)
        if dataset_name not in h5f:
            shape[0] = None
            h5f.create_dataset(dataset_name, data=data, maxshape=tuple(shape), chunks=True)
            return
        ds = h5f[dataset_name]
        old = ds.shape[0]
        shape[0] = old + data.shape[0]
        ds.resize(tuple(shape))
        ds[old:shape[0]] = data

    def ReadFeatures(self, h5_file_path):
        if str(h5_file_path).lower().endswith(".npy"):
            self.dataset = np.load(h5_file_path, mmap_mode="r")
            return
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("h5py is not installed: pass a .npy feature store or install h5py") from e
        self.h5py_file = h5py.File(h5_file_path, "r")
        self.dataset = self.h5py_file["dataset"]

    def GetFeaturesByID(self, idx):
        if idx >= len(self.dataset):
            raise IndexError("index error!")                       # (the reference's `raise("index error!")` is a TypeError)
        return np.asarray(self.dataset[idx])

    def Close(self):
        if self.h5py_file is not None:
            self.h5py_file.close()
            self.h5py_file = None


def score_layers(store, polygon_layer, line_layer):
    """test_for_shp's loop (ExtractFeatures.py:164-222) for EVERY line of `line_layer` (the reference `break`s
    after the first): pool each polygon's `PointID` rows of `store`, score every LEFT_FID / RIGHT_FID pair with
    the Euclidean distance and write it to the line attribute `simi` (created as OFTReal when missing).
    Layers are OGR layers or deepmerge_b200.shapefile.ShapefileLayer objects.  -> (fids, left, right, simi).
    Pooling and scoring run on the GPU (`pool_and_score`); only the attribute tables are host work."""
    from . import shapefile
    fids, left, right = [], [], []
    if isinstance(line_layer, shapefile.ShapefileLayer):            # whole columns at once
        t = line_layer.table
        l, r = t.column_int("LEFT_FID"), t.column_int("RIGHT_FID")
        keep = (l != -1) & (r != -1) & ~t.deleted
        fids, left, right = np.nonzero(keep)[0].tolist(), l[keep].tolist(), r[keep].tolist()
    else:
        line_layer.ResetReading()
        f = line_layer.GetNextFeature()
        while f is not None:
            a, b = int(f.GetField("LEFT_FID")), int(f.GetField("RIGHT_FID"))
            if a != -1 and b != -1:
                fids.append(int(f.GetFID()))
                left.append(a)
                right.append(b)
            f = line_layer.GetNextFeature()
    if isinstance(polygon_layer, shapefile.ShapefileLayer):
        point_ids = polygon_layer.table.column_str("PointID")
    else:
        point_ids = []
        for i in range(polygon_layer.GetFeatureCount()):
            v = polygon_layer.GetFeature(i).GetField("PointID")
            point_ids.append("" if v is None else str(v))
    _, simi = pool_and_score(np.asarray(store), point_ids, left, right)
    if line_layer.GetLayerDefn().GetFieldIndex("simi") < 0:
        try:
            from osgeo import ogr
            defn = ogr.FieldDefn("simi", ogr.OFTReal)
        except ImportError:
            defn = shapefile.FieldDefn("simi", shapefile.OFTReal)
        line_layer.CreateField(defn, 1)
    if isinstance(line_layer, shapefile.ShapefileLayer):
        line_layer.table.set_column("simi", np.asarray(simi, np.float64), rows=fids)
        line_layer.table.flush()
    else:
        for fid, v in zip(fids, simi):
            feat = line_layer.GetFeature(fid)
            feat.SetField("simi", float(v))
            line_layer.SetFeature(feat)
    return np.asarray(fids, np.int64), np.asarray(left, np.int64), np.asarray(right, np.int64), np.asarray(simi, np.float64)


def test_for_shp(featureIO, image_path=None, polygon_path=None, polyline_path=None, point_path=None):
    """The reference's test_for_shp(featureIO) (ExtractFeatures.py:150-225) with its hard-coded paths as
    arguments, every edge scored (no `break`), pooling + scoring on the GPU.  Returns 0 like the reference."""
    from .MyUtils2 import PolygonConnectPointDataset
    data_set = PolygonConnectPointDataset(image_path, polygon_path, polyline_path, point_path)
    store = featureIO.dataset if getattr(featureIO, "dataset", None) is not None else featureIO
    score_layers(store, data_set.polygon_layer, data_set.line_layer)
    return 0


test_for_shp.__test__ = False      # a reference-named entry point, not a pytest test
