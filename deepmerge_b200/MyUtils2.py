"""Drop-in for the RAG edge-list reader of the reference's MyUtils2.py.

PolygonConnectPointDataset (MyUtils2.py:128-209) iterates the polyline layer of lines.shp and
keeps [line_fid, tile_name, LEFT_FID, RIGHT_FID] for every feature whose two ids are not -1
(:184-186).  GDAL/OGR is not part of this image, so the class takes already-opened layer
objects (anything with ResetReading / GetNextFeature / GetField / GetFID, i.e. real OGR layers
or the in-memory stand-ins used by the tests) next to the reference's path arguments; with
osgeo importable it opens the paths exactly like the reference.  `edge_keys()` hands the list
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
        except ImportError as e:
            raise ValueError("Can not open {0}".format(self.polyline_path)) from e
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
