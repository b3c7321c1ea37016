"""On-disk adaptor for the shapefiles the reference reads and writes on the hot path (SURVEY.md
section 8(f) N3), without GDAL.

The reference goes through OGR for three things on this path:
  * `lines.shp` attributes LEFT_FID / RIGHT_FID -> the RAG edge list      (MyUtils2.py:155-193)
  * polygon attributes `PointID` (space separated), `join` / `Points` (comma separated) -> membership
    and neighbour lists                                                  (ExtractFeatures.py:175-179,
                                                                          MyUtils.py:110-117)
  * the score written back as the OFTReal attribute `simi` of the line   (ExtractFeatures.py:181-186,
                                                                          217-219)
and for the sample-point coordinates of `PointsGCS.shp` (MyUtils2.py:236-242).  All of it lives in
the .dbf attribute table (dBase III, fixed-width ASCII records) and, for the points, in fixed-size
.shp records -- both simple published formats, restated here in plain Python + numpy:

  DbfTable        the attribute table: vectorised column reads / writes (whole-column numpy parsing,
                  which is what feeds the GPU path), field creation, in-place record updates
  read_points     X / Y of a Point / PointZ / PointM .shp
  ShapefileLayer  the subset of the OGR layer / feature interface the reference calls on this path
                  (ResetReading, GetNextFeature, GetFeature, GetFeatureCount, GetLayerDefn().GetFieldIndex,
                  CreateField, SetFeature; feature.GetField / SetField / GetFID / GetGeometryRef().GetX/GetY),
                  so the reference-shaped classes (deepmerge_b200.MyUtils2.PolygonConnectPointDataset) run on
                  real files when `osgeo` is not installed
  write_dbf / write_point_shp   writers (tests, and exporting a raster-derived RAG for GIS tools)

Conventions follow the OGR "ESRI Shapefile" driver: FID = record index, trailing blanks of strings
are stripped, N fields without decimals are integers, blank / '*' filled numbers are null (None),
OFTReal is created as N(24,15), deleted records ('*' flag) are skipped by iteration.
"""
from __future__ import annotations

import datetime
import os
import struct

import numpy as np

OFTInteger, OFTReal, OFTString, OFTInteger64 = 0, 2, 4, 12      # the OGR field type codes the reference uses


class FieldDefn:
    """Name + dBase type ('C', 'N', 'F', 'L', 'D') + width + decimals of one attribute column."""

    def __init__(self, name, ftype="N", width=None, decimals=None):
        if isinstance(ftype, int):                                  # an OGR type code, as in ogr.FieldDefn(name, ogr.OFTReal)
            ftype, width, decimals = {OFTInteger: ("N", width or 9, 0), OFTInteger64: ("N", width or 18, 0),
                                      OFTReal: ("N", width or 24, 15 if decimals is None else decimals),
                                      OFTString: ("C", width or 80, 0)}[ftype]
        self.name, self.type = str(name), str(ftype).upper()
        self.width = int(width if width is not None else {"C": 80, "N": 24, "F": 24, "L": 1, "D": 8}[self.type])
        self.decimals = int(decimals or 0)
        if len(self.name.encode("ascii")) > 10:
            raise ValueError("dBase field names are at most 10 bytes: %r" % self.name)
        if not 1 <= self.width <= 254:
            raise ValueError("field width out of range: %d" % self.width)

    def GetName(self):
        return self.name

    def GetNameRef(self):
        return self.name

    @property
    def is_integer(self):
        return self.type == "N" and self.decimals == 0

    @property
    def is_real(self):
        return self.type == "F" or (self.type == "N" and self.decimals > 0)


class DbfTable:
    """A .dbf file held as one [n_records, record_length] byte matrix (numpy), parsed column-wise."""

    def __init__(self, path, update=False):
        self.path, self.update = path, bool(update)
        try:
            with open(path, "rb") as f:
                raw = f.read()
        except OSError as e:
            raise ValueError("Can not open {0}".format(path)) from e
        if len(raw) < 33:
            raise ValueError("Can not open {0}".format(path))
        self.version = raw[0]
        n, hlen, rlen = struct.unpack_from("<IHH", raw, 4)
        self.ldid = raw[29]
        self.fields = []
        off = 32
        while off + 32 <= hlen and raw[off] != 0x0D:
            name = raw[off:off + 11].split(b"\0", 1)[0].decode("ascii", "replace")
            ftype, width, dec = chr(raw[off + 11]), raw[off + 16], raw[off + 17]
            if ftype == "C":                                       # long character fields keep the high byte here
                width, dec = width + 256 * dec, 0
            fd = FieldDefn(name, ftype, max(1, min(width, 254)), dec)
            fd.width = width
            self.fields.append(fd)
            off += 32
        if rlen != 1 + sum(f.width for f in self.fields):
            raise ValueError("Can not open {0}".format(path))      # corrupt header
        avail = max(0, (len(raw) - hlen) // rlen) if rlen else 0
        n = min(n, avail)
        self.header_len, self.record_len = hlen, rlen
        self.records = np.frombuffer(raw, np.uint8, n * rlen, hlen).reshape(n, rlen).copy()
        self._dirty = False

    # ---- structure ----------------------------------------------------------------------------
    def __len__(self):
        return self.records.shape[0]

    def field_index(self, name):
        for i, f in enumerate(self.fields):
            if f.name.upper() == str(name).upper():                 # OGR matches field names case-insensitively
                return i
        return -1

    def _col(self, name_or_index):
        i = name_or_index if isinstance(name_or_index, int) else self.field_index(name_or_index)
        if i < 0 or i >= len(self.fields):
            raise KeyError(name_or_index)
        start = 1 + sum(f.width for f in self.fields[:i])
        return self.fields[i], start

    @property
    def deleted(self):
        """bool [n]: records carrying the dBase deletion flag."""
        return self.records[:, 0] == ord("*") if len(self) else np.zeros(0, bool)

    # ---- column reads (vectorised) ------------------------------------------------------------
    def column_bytes(self, name):
        """The fixed-width raw column as a numpy 'S<width>' array (padding included)."""
        f, s = self._col(name)
        return np.ascontiguousarray(self.records[:, s:s + f.width]).view("S%d" % f.width).reshape(-1)

    def column_str(self, name):
        """list[str] with trailing blanks stripped, as OGR returns string fields."""
        return [b.decode("utf-8", "replace").rstrip(" \0") for b in self.column_bytes(name).tolist()]

    def column_int(self, name, null=-1):
        """int64 [n] of a numeric column; blank / '*' (null) entries become `null`."""
        col = np.char.strip(self.column_bytes(name))
        bad = (col == b"") | (np.char.startswith(col, b"*"))
        out = np.full(col.shape, null, np.int64)
        if (~bad).any():
            good = col[~bad]
            try:
                out[~bad] = good.astype(np.int64)
            except ValueError:                                      # '12.000' style integers
                out[~bad] = good.astype(np.float64).astype(np.int64)
        return out

    def column_float(self, name, null=np.nan):
        col = np.char.strip(self.column_bytes(name))
        bad = (col == b"") | (np.char.startswith(col, b"*"))
        out = np.full(col.shape, null, np.float64)
        if (~bad).any():
            out[~bad] = col[~bad].astype(np.float64)
        return out

    def value(self, rec, name_or_index):
        """One cell with OGR's typing: int / float / str / None (null)."""
        f, s = self._col(name_or_index)
        raw = bytes(self.records[rec, s:s + f.width])
        if f.type == "C":
            return raw.decode("utf-8", "replace").rstrip(" \0")
        txt = raw.strip(b" \0")
        if txt == b"" or txt.startswith(b"*"):
            return None
        if f.type in ("N", "F"):
            if f.is_integer:
                try:
                    return int(txt)
                except ValueError:
                    return int(float(txt))
            return float(txt)
        if f.type == "L":
            return txt[:1] in b"YyTt"
        return txt.decode("ascii", "replace")                      # 'D' and anything else: the text

    # ---- writes -------------------------------------------------------------------------------
    @staticmethod
    def _format(f, v):
        if v is None:
            return b" " * f.width
        if f.type == "C":
            b = str(v).encode("utf-8")[:f.width]
            return b + b" " * (f.width - len(b))
        if f.type in ("N", "F"):
            txt = ("%d" % int(v)) if f.is_integer else ("%.*f" % (f.decimals, float(v)))
            if len(txt) > f.width:                                  # OGR truncates decimals before giving up
                txt = ("%.*g" % (max(1, f.width - 7), float(v)))[:f.width]
            return txt.rjust(f.width).encode("ascii")
        if f.type == "L":
            return (b"T" if v else b"F").ljust(f.width)
        return str(v).encode("ascii")[:f.width].ljust(f.width)

    def set_value(self, rec, name_or_index, v):
        f, s = self._col(name_or_index)
        self.records[rec, s:s + f.width] = np.frombuffer(self._format(f, v), np.uint8)
        self._dirty = True

    def set_column(self, name, values, rows=None):
        """Write a whole column (or the records `rows`) from a sequence / numpy array -- vectorised for numbers."""
        f, s = self._col(name)
        rows = np.arange(len(self)) if rows is None else np.asarray(rows, np.int64)
        values = np.asarray(values)
        if values.shape[0] != rows.shape[0]:
            raise ValueError("one value per record expected")
        if rows.size == 0:
            return
        if f.type in ("N", "F") and values.dtype.kind in "iuf":
            if f.is_integer:
                txt = np.char.mod("%d", values.astype(np.int64))
            else:
                txt = np.char.mod("%%.%df" % f.decimals, values.astype(np.float64))
            txt = np.char.rjust(txt, f.width).astype("S")
            if txt.dtype.itemsize != f.width:
                raise ValueError("a value does not fit field %s N(%d,%d)" % (f.name, f.width, f.decimals))
            self.records[rows, s:s + f.width] = txt.view(np.uint8).reshape(-1, f.width)
        else:
            for r, v in zip(rows.tolist(), values.tolist()):
                self.records[r, s:s + f.width] = np.frombuffer(self._format(f, v), np.uint8)
        self._dirty = True

    def add_field(self, defn: FieldDefn):
        """Append a column (all null), as OGR's CreateField does on a shapefile layer."""
        if self.field_index(defn.name) >= 0:
            raise ValueError("field %s exists" % defn.name)
        self.fields.append(defn)
        pad = np.full((len(self), defn.width), ord(" "), np.uint8)
        self.records = np.concatenate([self.records, pad], axis=1)
        self.record_len += defn.width
        self.header_len = 32 + 32 * len(self.fields) + 1
        self._dirty = True

    def flush(self):
        if not self._dirty:
            return
        if not self.update:
            raise ValueError("{0} was opened read-only".format(self.path))
        write_dbf(self.path, self.fields, records=self.records, ldid=self.ldid)
        self._dirty = False


def _header_bytes(fields, n_records, ldid=0):
    today = datetime.date.today()
    rlen = 1 + sum(f.width for f in fields)
    hlen = 32 + 32 * len(fields) + 1
    head = bytearray(32)
    head[0] = 0x03
    head[1:4] = bytes([today.year - 1900, today.month, today.day])
    struct.pack_into("<IHH", head, 4, n_records, hlen, rlen)
    head[29] = ldid
    out = bytes(head)
    for f in fields:
        d = bytearray(32)
        nm = f.name.encode("ascii")
        d[:len(nm)] = nm
        d[11] = ord(f.type)
        d[16] = f.width & 0xFF
        d[17] = (f.width >> 8) if f.type == "C" else f.decimals
        out += bytes(d)
    return out + b"\x0D"


def write_dbf(path, fields, columns=None, records=None, ldid=0):
    """Write a dBase III table.  Either `columns` = {field name: sequence} or a ready
    [n, record_length] uint8 `records` matrix."""
    fields = [f if isinstance(f, FieldDefn) else FieldDefn(*f) for f in fields]
    if records is None:
        n = len(next(iter(columns.values()))) if columns else 0
        rlen = 1 + sum(f.width for f in fields)
        records = np.full((n, rlen), ord(" "), np.uint8)
        off = 1
        for f in fields:
            vals = columns[f.name]
            if len(vals) != n:
                raise ValueError("column %s has %d values, expected %d" % (f.name, len(vals), n))
            for r, v in enumerate(vals):
                records[r, off:off + f.width] = np.frombuffer(DbfTable._format(f, v), np.uint8)
            off += f.width
    tmp = path + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(_header_bytes(fields, records.shape[0], ldid))
        fh.write(np.ascontiguousarray(records).tobytes())
        fh.write(b"\x1A")
    os.replace(tmp, path)


# ---------------------------------------------------------------------------------------------
# .shp geometry (points)
# ---------------------------------------------------------------------------------------------
_POINT_TYPES = (1, 11, 21)          # Point, PointZ, PointM


def _shp_header(path):
    try:
        with open(path, "rb") as f:
            raw = f.read()
    except OSError as e:
        raise ValueError("Can not open {0}".format(path)) from e
    if len(raw) < 100 or struct.unpack_from(">i", raw, 0)[0] != 9994:
        raise ValueError("Can not open {0}".format(path))
    return raw, struct.unpack_from("<i", raw, 32)[0]


def read_points(path):
    """-> (X float64 [n], Y float64 [n]) of a point shapefile; null shapes give NaN."""
    raw, shape_type = _shp_header(path)
    if shape_type not in _POINT_TYPES:
        raise ValueError("{0} is not a point shapefile (shape type {1})".format(path, shape_type))
    xs, ys = [], []
    off, end = 100, len(raw)
    while off + 8 <= end:
        (clen,) = struct.unpack_from(">i", raw, off + 4)
        body = off + 8
        (st,) = struct.unpack_from("<i", raw, body) if body + 4 <= end else (0,)
        if st in _POINT_TYPES and body + 20 <= end:
            x, y = struct.unpack_from("<dd", raw, body + 4)
        else:
            x = y = float("nan")
        xs.append(x)
        ys.append(y)
        off = body + 2 * clen
    return np.asarray(xs, np.float64), np.asarray(ys, np.float64)


def shape_type(path):
    return _shp_header(path)[1]


def write_point_shp(path, X, Y):
    """Point .shp + .shx for coordinates X, Y (tests; exporting sample points)."""
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    n = X.shape[0]
    bbox = (float(X.min()), float(Y.min()), float(X.max()), float(Y.max())) if n else (0.0, 0.0, 0.0, 0.0)

    def header(file_words):
        h = bytearray(100)
        struct.pack_into(">i", h, 0, 9994)
        struct.pack_into(">i", h, 24, file_words)
        struct.pack_into("<ii", h, 28, 1000, 1)
        struct.pack_into("<dddd", h, 36, *bbox)
        return bytes(h)

    recs, idx = bytearray(), bytearray()
    for i in range(n):
        idx += struct.pack(">ii", (100 + len(recs)) // 2, 10)
        recs += struct.pack(">ii", i + 1, 10) + struct.pack("<idd", 1, X[i], Y[i])
    base = path[:-4] if path.lower().endswith(".shp") else path
    with open(base + ".shp", "wb") as f:
        f.write(header((100 + len(recs)) // 2) + bytes(recs))
    with open(base + ".shx", "wb") as f:
        f.write(header((100 + len(idx)) // 2) + bytes(idx))


# ---------------------------------------------------------------------------------------------
# the OGR-shaped view the reference's code paths call
# ---------------------------------------------------------------------------------------------
class _Point:
    def __init__(self, x, y):
        self._x, self._y = float(x), float(y)

    def GetX(self, i=0):
        return self._x

    def GetY(self, i=0):
        return self._y


class Feature:
    def __init__(self, layer, fid):
        self._layer, self._fid, self._pending = layer, int(fid), {}

    def GetFID(self):
        return self._fid

    def GetField(self, name_or_index):
        i = name_or_index if isinstance(name_or_index, int) else self._layer.table.field_index(name_or_index)
        if i < 0:
            raise KeyError("no field %r" % (name_or_index,))       # OGR raises too (KeyError since GDAL 3.x)
        if i in self._pending:
            return self._pending[i]
        return self._layer.table.value(self._fid, i)

    def GetFieldAsString(self, name_or_index):
        v = self.GetField(name_or_index)
        return "" if v is None else str(v)

    def SetField(self, name_or_index, value):
        i = name_or_index if isinstance(name_or_index, int) else self._layer.table.field_index(name_or_index)
        if i < 0:
            raise KeyError("no field %r" % (name_or_index,))
        self._pending[i] = value

    def GetGeometryRef(self):
        xy = self._layer.points
        if xy is None or self._fid >= xy[0].shape[0] or np.isnan(xy[0][self._fid]):
            return None
        return _Point(xy[0][self._fid], xy[1][self._fid])


class _LayerDefn:
    def __init__(self, table):
        self._t = table

    def GetFieldIndex(self, name):
        return self._t.field_index(name)

    def GetFieldCount(self):
        return len(self._t.fields)

    def GetFieldDefn(self, i):
        return self._t.fields[i]


class ShapefileLayer:
    """`drv.Open(path, update).GetLayer(0)` of the reference, for the calls made on this path."""

    def __init__(self, path, update=0):
        base = path[:-4] if path.lower().endswith((".shp", ".dbf")) else path
        self.path = path
        self.table = DbfTable(base + ".dbf", update=bool(update))
        self._shp = base + ".shp"
        self._points = False            # not loaded yet
        self._cursor = 0

    @property
    def points(self):
        if self._points is False:
            self._points = None
            if os.path.exists(self._shp) and shape_type(self._shp) in _POINT_TYPES:
                self._points = read_points(self._shp)
        return self._points

    def GetLayerDefn(self):
        return _LayerDefn(self.table)

    def GetFeatureCount(self, force=1):
        return int((~self.table.deleted).sum())

    def ResetReading(self):
        self._cursor = 0

    def GetNextFeature(self):
        n, dele = len(self.table), self.table.deleted
        while self._cursor < n and dele[self._cursor]:
            self._cursor += 1
        if self._cursor >= n:
            return None
        self._cursor += 1
        return Feature(self, self._cursor - 1)

    def __iter__(self):
        self.ResetReading()
        while True:
            f = self.GetNextFeature()
            if f is None:
                return
            yield f

    def GetFeature(self, fid):
        fid = int(fid)
        if fid < 0 or fid >= len(self.table) or self.table.deleted[fid]:
            return None                                             # OGR returns None for a missing FID
        return Feature(self, fid)

    def CreateField(self, defn, approx_ok=1):
        if not isinstance(defn, FieldDefn):                         # a real ogr.FieldDefn
            defn = FieldDefn(defn.GetName(), defn.GetType(), defn.GetWidth() or None, defn.GetPrecision() or None)
        if not self.table.update:
            return 6                                                # OGRERR_FAILURE: layer not opened for update
        self.table.add_field(defn)
        self.table.flush()
        return 0

    def SetFeature(self, feature):
        if not self.table.update:
            return 6
        for i, v in feature._pending.items():
            self.table.set_value(feature.GetFID(), i, v)
        feature._pending.clear()
        self.table.flush()
        return 0

    def SyncToDisk(self):
        self.table.flush()
        return 0


class ShapefileDataSource:
    """`ogr.GetDriverByName("ESRI Shapefile").Open(path, update)`: one layer."""

    def __init__(self, path, update=0):
        self._layer = ShapefileLayer(path, update)

    def GetLayer(self, i=0):
        return self._layer if i == 0 else None

    def GetLayerCount(self):
        return 1


def Open(path, update=0):
    """None when the file cannot be opened, like the OGR driver (callers raise "Can not open ...")."""
    try:
        return ShapefileDataSource(path, update)
    except ValueError:
        return None
