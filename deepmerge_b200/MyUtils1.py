"""Drop-in for the adjacent-pair sampler of the reference's MyUtils1.py (R10).

MergingSegmensPairDataset.add_data (MyUtils1.py:236-295): for every labelled polygon pair of a
pair-list txt (columns 1,2 = left,right polygon id, :225-234) pick ONE member sample point on each
side with two `random.randint` draws in that order (:278-279), and emit [tile, left_pt, right_pt,
flag].  The same draws in the same order reproduce the reference's list under the same seed.
`sample_pairs` is the array form used by the GPU training step (dm_gather_rows + Loss);
`MergingSegmensPairDataset` is the class itself: the reference's constructor and item layout (:20-58)."""
from __future__ import annotations

import os
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


def _open_vector(path):
    """OGR when it is installed, else the GDAL-free adaptor; same failure as the reference (MyUtils1.py:307-315)."""
    try:
        from osgeo import ogr
        ds = ogr.GetDriverByName("ESRI Shapefile").Open(path, 0)
    except ImportError:
        from . import shapefile
        ds = shapefile.Open(path, 0)
    if ds is None or ds.GetLayer(0) is None:
        raise ValueError("Can not open {0}".format(path))
    return ds, ds.GetLayer(0)


def _open_image(path):
    try:
        from osgeo import gdal
        ds = gdal.Open(path, gdal.GA_ReadOnly)
    except ImportError:
        from . import geotiff
        ds = geotiff.Open(path)
    if ds is None:
        raise ValueError("Can not open {0}".format(path))
    return ds


class MergingSegmensPairDataset:
    """Drop-in for the reference's training-pair dataset (MyUtils1.py:20-58): same constructor, same attributes
    (`data`, `layers`, `img_dataset`, `positive_number`, `positive_pair_number`, `negative_number`,
    `negative_pair_number`), `len()`, and `ds[i] -> (left_meta, right_meta, flag)` with
    `meta = (designed [1, 19] tensor, scales [1, 4] tensor, [4 float32 arrays [C, s, s]])`.

    * the pair list is sampled exactly as add_data does (:236-295): one uniformly random member point per side with
      two `random.randint` draws per labelled pair, in that order -- the same seed gives the same list;
    * the patches of an item (get_patches_by_scales :117-127: window, zero-padded cut, INTER_AREA resize, / 255) come from
      the CUDA loader kernels (dm_cut_windows, dm_resize_area), bit-identical to the reference's cv2 path;
    * `batch(indices)` is the form a training step at scale wants: everything for B pairs at once, on the device.
    `open_vector(path) -> (datasource, layer)` and `open_image(path) -> dataset` default to GDAL/OGR when installed, else
    to deepmerge_b200.shapefile / deepmerge_b200.geotiff."""

    def __init__(self, image_folder, polygon_folder, point_folder, positive_folder, negative_folder, num=0, *,
                 open_vector=None, open_image=None, device=None):
        self.image_folder, self.polygon_folder, self.point_folder = image_folder, polygon_folder, point_folder
        self.num = num
        self.positive_folder, self.negative_folder = positive_folder, negative_folder
        self._open_vector, self._open_image = open_vector or _open_vector, open_image or _open_image
        self._device = device
        self.data = []
        self.point_dataset = []          # the reference keeps the data sources alive here (:29)
        self.img_dataset = {}
        self.layers = {}
        self._image_dev = {}
        self.positive_number, self.positive_pair_number = self.add_data(self.get_all_files(positive_folder), 1)
        self.negative_number, self.negative_pair_number = self.add_data(self.get_all_files(negative_folder), 0)

    def __len__(self):
        return len(self.data)

    def get_all_files(self, cwd):
        """MyUtils1.py:296-305."""
        if cwd == "" or cwd is None:
            return None
        return [os.path.join(cwd, name) for name in os.listdir(cwd)]

    def add_data(self, txt_path, flag):
        """MyUtils1.py:236-295 -> (items added, labelled pairs read)."""
        if txt_path is None:
            return 0, 0
        count = pair_count = 0
        for path in txt_path:
            name = os.path.basename(str(path).replace("\\", "/")).split(".")[0]
            _poly_ds, polygon_layer = self._open_vector(os.path.join(self.polygon_folder, name + ".shp"))
            _pt_ds, point_layer = self._open_vector(os.path.join(self.point_folder, name, "PointsGCS.shp"))
            img = self._open_image(os.path.join(self.image_folder, name + ".tif"))
            self.point_dataset.append(_pt_ds)
            self.img_dataset[name] = img
            self.band_num = img.RasterCount
            self.layers[name] = point_layer
            pairs = read_pair_list(path)

            class _Fields:                      # point_id_fields[i] = PointID of polygon i, read when it is needed
                def __getitem__(_, i):
                    return polygon_layer.GetFeature(int(i)).GetField("PointID")

            items = sample_pairs(_Fields(), pairs, flag, name)
            self.data.extend(items)
            count += len(items)
            pair_count += len(pairs)
        return count, pair_count

    # ---- items ------------------------------------------------------------------------------------------------
    def _device_image(self, tile):
        import torch
        if tile not in self._image_dev:
            ds = self.img_dataset[tile]
            arr = np.ascontiguousarray(ds.ReadAsArray(0, 0, ds.RasterXSize, ds.RasterYSize), dtype=np.uint8)
            dev = self._device if self._device is not None else torch.device("cuda", torch.cuda.current_device())
            self._image_dev[tile] = torch.from_numpy(arr).to(dev)
        return self._image_dev[tile]

    def _windows(self, tile, point_ids):
        """designed [n, 19] float32 (15 attributes + 4 scale factors), scales int64 [n, 4], pixel positions."""
        from . import MyUtils2 as m2
        layer, ds = self.layers[tile], self.img_dataset[tile]
        feats = [layer.GetFeature(int(i)) for i in point_ids]
        attr = np.asarray([[float(f.GetField(n)) for n in m2.DESIGNED_FIELDS] for f in feats], np.float64).reshape(-1, 15)
        inner = np.asarray([int(f.GetField("inner")) for f in feats], np.int64)
        obj = np.asarray([int(f.GetField("object")) for f in feats], np.int64)
        interval = obj - inner                                       # get_scales, MyUtils1.py:129-152
        scales = np.stack([inner, obj, obj + interval, obj + 2 * interval], axis=1).astype(np.int64)
        factors = scales / np.asarray(m2.SCALES, np.float64)
        X = np.asarray([f.GetGeometryRef().GetX() for f in feats], np.float64)
        Y = np.asarray([f.GetGeometryRef().GetY() for f in feats], np.float64)
        xpix, ylin = m2.geo_to_pixel(ds.GetGeoTransform(), X, Y)     # the "+ 1" convention of :73-74
        return {"ids": np.asarray(point_ids, np.int64), "designed": np.concatenate([attr, factors], axis=1).astype(np.float32),
                "scales": scales, "xpix": xpix, "ylin": ylin}

    def get_all_features(self, tile, point_id):
        """MyUtils1.py:60-80 for one sample point -> (designed [1,19] tensor, scales [1,4] tensor, [4 patches])."""
        import torch
        from . import MyUtils2 as m2
        w = self._windows(tile, [int(point_id)])
        patches = m2.point_patches(self._device_image(tile), w)
        return (torch.from_numpy(w["designed"]), torch.from_numpy(w["scales"].astype(np.float32)),
                [p[0].cpu().numpy() for p in patches])

    def __getitem__(self, index):
        tile, left, right, flag = self.data[index]
        return self.get_all_features(tile, left), self.get_all_features(tile, right), int(flag)

    def batch(self, indices):
        """B items at once on the device -> dict(left=(designed [B,19], scales [B,4], [4 x [B,C,s,s]]), right=..., flag
        int64 [B]): one cut + resize launch group per tile and window size instead of 8 B of them."""
        import torch
        from . import MyUtils2 as m2
        idx = [int(i) for i in indices]
        B = len(idx)
        dev = None
        out = {}
        for side, col in (("left", 1), ("right", 2)):
            designed = scales = None
            patches = None
            by_tile = {}
            for b, i in enumerate(idx):
                by_tile.setdefault(self.data[i][0], []).append(b)
            for tile, rows in by_tile.items():
                w = self._windows(tile, [int(self.data[idx[b]][col]) for b in rows])
                img = self._device_image(tile)
                dev = img.device
                pp = m2.point_patches(img, w)
                if patches is None:
                    designed = torch.empty((B, 19), dtype=torch.float32, device=dev)
                    scales = torch.empty((B, 4), dtype=torch.float32, device=dev)
                    patches = [torch.empty((B,) + tuple(p.shape[1:]), dtype=torch.float32, device=dev) for p in pp]
                r = torch.as_tensor(rows, device=dev)
                designed[r] = torch.from_numpy(w["designed"]).to(dev)
                scales[r] = torch.from_numpy(w["scales"].astype(np.float32)).to(dev)
                for k in range(4):
                    patches[k][r] = pp[k]
            out[side] = (designed, scales, patches)
        out["flag"] = torch.as_tensor([int(self.data[i][3]) for i in idx], dtype=torch.int64, device=dev)
        return out
