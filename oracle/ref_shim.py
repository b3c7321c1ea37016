"""Import the reference IN PLACE from /root/reference (TEST INFRASTRUCTURE, build
container only; the GPU box has no /root/reference and nothing at run time reads it).

Recipe from SURVEY.md Appendix B: three sys.modules shims (timm.models.layers, osgeo,
h5py) are enough for ExtractFeatures / MyUtils1 / MyUtils2 / Nets / Losses to import.
The fake OGR/GDAL objects below drive the reference's dataset classes without GDAL.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE = "/root/reference"


def available():
    return os.path.isdir(REFERENCE)


def load():
    """Returns a namespace with the reference modules that import here."""
    import torch.nn as nn

    sys.dont_write_bytecode = True          # /root/reference is read-only
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import vit_model                        # noqa: imports as-is

    tl = types.ModuleType("timm.models.layers")
    tl.trunc_normal_, tl.DropPath = nn.init.trunc_normal_, vit_model.DropPath
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    sys.modules.setdefault("timm.models", types.ModuleType("timm.models"))
    sys.modules["timm.models.layers"] = tl
    if "osgeo" not in sys.modules:
        osgeo = types.ModuleType("osgeo")
        osgeo.gdal = types.SimpleNamespace(GA_ReadOnly=0)
        osgeo.ogr = types.SimpleNamespace()
        sys.modules["osgeo"] = osgeo
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import ExtractFeatures, MyUtils1, MyUtils2, Nets, Losses, config   # noqa

    return types.SimpleNamespace(ExtractFeatures=ExtractFeatures, MyUtils1=MyUtils1, MyUtils2=MyUtils2,
                                 Nets=Nets, Losses=Losses, config=config, osgeo=sys.modules["osgeo"])


# ----------------------------- duck-typed OGR / GDAL fakes ----------------------------- #


class FakeGeometry:
    def __init__(self, x, y):
        self._x, self._y = x, y

    def GetX(self):
        return self._x

    def GetY(self):
        return self._y


class FakeFeature:
    def __init__(self, fid, fields, xy=None):
        self._fid, self._fields, self._xy = fid, dict(fields), xy

    def GetFID(self):
        return self._fid

    def GetField(self, name):
        return self._fields[name]

    def SetField(self, name, value):
        self._fields[name] = value

    def GetGeometryRef(self):
        return FakeGeometry(*self._xy)


class FakeLayer:
    def __init__(self, features):
        self._f = list(features)
        self._i = 0
        self.fields = set()

    def ResetReading(self):
        self._i = 0

    def GetNextFeature(self):
        if self._i >= len(self._f):
            return None
        self._i += 1
        return self._f[self._i - 1]

    def GetFeature(self, i):
        return self._f[int(i)]

    def SetFeature(self, f):
        self._f[f.GetFID()] = f


class FakeVectorDS:
    def __init__(self, layer):
        self._layer = layer

    def GetLayer(self, i):
        return self._layer


class FakeRaster:
    """numpy-backed stand-in for a GDAL dataset; arr is [C,H,W] uint8."""

    def __init__(self, arr, geotransform=(0.0, 1.0, 0.0, 0.0, 0.0, -1.0)):
        self.arr = arr
        self.RasterCount, self.RasterYSize, self.RasterXSize = arr.shape
        self._gt = geotransform

    def ReadAsArray(self, x, y, w, h):
        return self.arr[:, y:y + h, x:x + w]

    def GetGeoTransform(self):
        return self._gt


def install_fake_drivers(ref, layers_by_path, rasters_by_path):
    """Point the osgeo shim's ogr.GetDriverByName(...).Open / gdal.Open at in-memory fakes
    (MyUtils2.py:195-209 call exactly these)."""

    class Driver:
        def Open(self, path, mode=0):
            layer = layers_by_path.get(path)
            return None if layer is None else FakeVectorDS(layer)

    ref.osgeo.ogr.GetDriverByName = lambda name: Driver()
    ref.osgeo.gdal.Open = lambda path, mode=0: rasters_by_path.get(path)
    ref.osgeo.gdal.GA_ReadOnly = 0
