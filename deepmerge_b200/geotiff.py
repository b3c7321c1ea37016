"""On-disk adaptor for the scene rasters of the reference, without GDAL.

The reference opens its images with `gdal.Open(path)` and calls, on this path, only
`RasterCount / RasterXSize / RasterYSize`, `GetGeoTransform()` (sample point -> pixel,
MyUtils2.py:239-242) and `ReadAsArray(xoff, yoff, xsize, ysize)` (window cut, :330-360).
`Open(path)` returns an object with exactly those members for

  * TIFF / GeoTIFF files: classic and BigTIFF, strips or tiles, 8-bit samples, chunky or planar,
    uncompressed or Deflate (decoded here); anything else (LZW, predictors, 16-bit ...) is handed to
    Pillow when it is installed.  The geotransform comes from the GeoTIFF tags ModelPixelScale +
    ModelTiepoint (or ModelTransformation), with GDAL's half-pixel shift for PixelIsPoint rasters,
    else from a world file (.tfw / .tifw / .wld) beside the image, else GDAL's default (0,1,0,0,0,1);
  * `.npy` arrays [H, W] / [H, W, C] with an optional world file.

The whole raster is decoded into one [C, H, W] uint8 array (the layout ReadAsArray returns); at the
sizes of the north-star scenes (10k x 10k x 4 = 400 MB) that is also the host buffer the GPU path
uploads.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}


class RasterDataset:
    """What `gdal.Open` returns, for the members the reference uses on this path."""

    def __init__(self, array_chw, geotransform=(0.0, 1.0, 0.0, 0.0, 0.0, 1.0)):
        a = np.asarray(array_chw)
        if a.ndim == 2:
            a = a[None]
        if a.ndim != 3:
            raise ValueError("raster must be [C, H, W]")
        self.array = a
        self.RasterCount, self.RasterYSize, self.RasterXSize = int(a.shape[0]), int(a.shape[1]), int(a.shape[2])
        self._gt = tuple(float(v) for v in geotransform)

    def GetGeoTransform(self):
        return self._gt

    def ReadAsArray(self, xoff=0, yoff=0, xsize=None, ysize=None):
        xsize = self.RasterXSize - xoff if xsize is None else xsize
        ysize = self.RasterYSize - yoff if ysize is None else ysize
        if xoff < 0 or yoff < 0 or xsize < 0 or ysize < 0 or xoff + xsize > self.RasterXSize or yoff + ysize > self.RasterYSize:
            return None                                             # GDAL reports an error and returns None
        w = self.array[:, yoff:yoff + ysize, xoff:xoff + xsize]
        return w[0] if self.RasterCount == 1 else w

    def interleaved(self):
        """[H, W, C] uint8 copy: the layout the raster kernels take."""
        return np.ascontiguousarray(np.moveaxis(self.array, 0, 2))


# ---------------------------------------------------------------------------------------------
# TIFF
# ---------------------------------------------------------------------------------------------
def _read_ifd(raw, path):
    if len(raw) < 8 or raw[:2] not in (b"II", b"MM"):
        raise ValueError("Can not open {0}".format(path))
    e = "<" if raw[:2] == b"II" else ">"
    (magic,) = struct.unpack_from(e + "H", raw, 2)
    if magic == 42:
        big, (off,) = False, struct.unpack_from(e + "I", raw, 4)
        cnt_fmt, ent_size, val_size, n_fmt = "H", 12, 4, "I"
    elif magic == 43:
        big, (off,) = True, struct.unpack_from(e + "Q", raw, 8)
        cnt_fmt, ent_size, val_size, n_fmt = "Q", 20, 8, "Q"
    else:
        raise ValueError("Can not open {0}".format(path))
    (n,) = struct.unpack_from(e + cnt_fmt, raw, off)
    p = off + struct.calcsize(cnt_fmt)
    tags = {}
    for _ in range(n):
        tag, typ = struct.unpack_from(e + "HH", raw, p)
        (count,) = struct.unpack_from(e + n_fmt, raw, p + 4)
        fmt = _TYPE_FMT.get(typ)
        vpos = p + 4 + struct.calcsize(n_fmt)
        if fmt is not None:
            size = struct.calcsize("=" + fmt) * count
            if size > val_size:
                (vpos,) = struct.unpack_from(e + n_fmt, raw, vpos)
            if typ == 2:
                tags[tag] = raw[vpos:vpos + count].split(b"\0", 1)[0].decode("latin1")
            elif vpos + size <= len(raw):
                vals = struct.unpack_from(e + fmt * count, raw, vpos)
                if typ in (5, 10):
                    vals = tuple(vals[i] / vals[i + 1] if vals[i + 1] else 0.0 for i in range(0, len(vals), 2))
                tags[tag] = vals
        p += ent_size
    return tags


def _geotransform_from_tags(tags):
    scale, tie, mat = tags.get(33550), tags.get(33922), tags.get(34264)
    gt = None
    if mat and len(mat) >= 16:
        gt = (mat[3], mat[0], mat[1], mat[7], mat[4], mat[5])
    elif scale and tie and len(scale) >= 2 and len(tie) >= 6:
        sx, sy = scale[0], scale[1]
        gt = (tie[3] - tie[0] * sx, sx, 0.0, tie[4] + tie[1] * sy, 0.0, -sy)
    if gt is not None:
        keys = tags.get(34735)
        if keys and len(keys) >= 4:                                 # GeoKeyDirectory: (id, location, count, value) x n
            for i in range(4, 4 + 4 * keys[3], 4):
                if i + 3 < len(keys) and keys[i] == 1025 and keys[i + 1] == 0 and keys[i + 3] == 2:   # RasterPixelIsPoint
                    gt = (gt[0] - 0.5 * gt[1] - 0.5 * gt[2], gt[1], gt[2], gt[3] - 0.5 * gt[4] - 0.5 * gt[5], gt[4], gt[5])
    return gt


def _world_file(path):
    base, ext = os.path.splitext(path)
    cands = [base + ".tfw", base + ".tifw", base + ".wld", path + "w"]
    if len(ext) == 4:
        cands.insert(0, base + "." + ext[1] + ext[3] + "w")         # .tif -> .tfw, .png -> .pgw
    for c in cands:
        if os.path.exists(c):
            with open(c) as f:
                v = [float(t) for t in f.read().split()]
            if len(v) >= 6:
                A, D, B, E, C, F = v[:6]                             # centre of the upper-left pixel -> its corner
                return (C - 0.5 * A - 0.5 * B, A, B, F - 0.5 * D - 0.5 * E, D, E)
    return None


def _decode_tiff(raw, tags, path):
    W, H = tags[256][0], tags[257][0]
    spp = tags.get(277, (1,))[0]
    bits = tags.get(258, (1,))
    comp = tags.get(259, (1,))[0]
    planar = tags.get(284, (1,))[0]
    predictor = tags.get(317, (1,))[0]
    if any(b != 8 for b in bits) or comp not in (1, 8, 32946) or predictor != 1 or tags.get(339, (1,))[0] not in (1,):
        return None                                                 # not handled here
    tiled = 322 in tags
    offs = tags[324] if tiled else tags[273]
    cnts = tags[325] if tiled else tags[279]
    bw = tags[322][0] if tiled else W
    bh = tags[323][0] if tiled else min(tags.get(278, (H,))[0], H)
    nx, ny = -(-W // bw), -(-H // bh)
    planes = spp if planar == 2 else 1
    ch = 1 if planar == 2 else spp
    if len(offs) < nx * ny * planes:
        raise ValueError("Can not open {0}".format(path))
    out = np.zeros((spp, H, W), np.uint8)
    for pl in range(planes):
        for by in range(ny):
            for bx in range(nx):
                i = (pl * ny + by) * nx + bx
                buf = raw[offs[i]:offs[i] + cnts[i]]
                if comp != 1:
                    buf = zlib.decompress(buf)
                rows = bh if tiled else min(bh, H - by * bh)
                blk = np.frombuffer(buf, np.uint8, rows * bw * ch).reshape(rows, bw, ch)
                h, w = min(rows, H - by * bh), min(bw, W - bx * bw)
                dst = out[pl:pl + 1] if planar == 2 else out
                dst[:, by * bh:by * bh + h, bx * bw:bx * bw + w] = np.moveaxis(blk[:h, :w], 2, 0)
    return out


def _open_tiff(path):
    with open(path, "rb") as f:
        raw = f.read()
    tags = _read_ifd(raw, path)
    if 256 not in tags or 257 not in tags:
        raise ValueError("Can not open {0}".format(path))
    arr = _decode_tiff(raw, tags, path)
    if arr is None:
        try:
            from PIL import Image
        except ImportError as e:
            raise ValueError("Can not open {0}: TIFF variant needs Pillow".format(path)) from e
        Image.MAX_IMAGE_PIXELS = None
        with Image.open(path) as im:
            a = np.asarray(im)
        if a.dtype != np.uint8:
            raise ValueError("Can not open {0}: only 8-bit rasters are supported".format(path))
        arr = a[None] if a.ndim == 2 else np.moveaxis(a, 2, 0)
    gt = _geotransform_from_tags(tags) or _world_file(path) or (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    return RasterDataset(arr, gt)


def Open(path, mode=0):
    """`gdal.Open(path, gdal.GA_ReadOnly)`: a RasterDataset, or None when the file cannot be opened."""
    try:
        if str(path).lower().endswith(".npy"):
            a = np.load(path)
            if a.dtype != np.uint8 or a.ndim not in (2, 3):
                return None
            return RasterDataset(a if a.ndim == 2 else np.moveaxis(a, 2, 0), _world_file(path) or (0.0, 1.0, 0.0, 0.0, 0.0, 1.0))
        return _open_tiff(path)
    except (OSError, ValueError, KeyError, struct.error, zlib.error):
        return None


def write_geotiff(path, array_hwc, geotransform=None, tile=None, deflate=False):
    """Baseline (Geo)TIFF writer: 8-bit, chunky, strips of 64 rows or square tiles, optional Deflate.
    For tests and for exporting label-derived rasters to GIS tools."""
    a = np.asarray(array_hwc, np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    H, W, C = a.shape
    blocks = []
    if tile:
        for y in range(0, H, tile):
            for x in range(0, W, tile):
                blk = np.zeros((tile, tile, C), np.uint8)
                sub = a[y:y + tile, x:x + tile]
                blk[:sub.shape[0], :sub.shape[1]] = sub
                blocks.append(blk.tobytes())
    else:
        rps = 64
        for y in range(0, H, rps):
            blocks.append(np.ascontiguousarray(a[y:y + rps]).tobytes())
    if deflate:
        blocks = [zlib.compress(b, 6) for b in blocks]
    entries = []          # (tag, type, values)

    def add(tag, typ, vals):
        entries.append((tag, typ, tuple(vals)))

    add(256, 4, [W]); add(257, 4, [H]); add(258, 3, [8] * C); add(259, 3, [8 if deflate else 1])
    add(262, 3, [2 if C >= 3 else 1]); add(277, 3, [C]); add(284, 3, [1])
    if C > 3:
        add(338, 3, [2] + [0] * (C - 4))                            # 4th sample as unassociated alpha: generic readers keep it
    if tile:
        add(322, 4, [tile]); add(323, 4, [tile])
    else:
        add(278, 4, [64])
    if geotransform is not None:
        gt = geotransform
        add(33550, 12, [gt[1], -gt[5], 0.0]); add(33922, 12, [0.0, 0.0, 0.0, gt[0], gt[3], 0.0])
        add(34735, 3, [1, 1, 0, 1, 1025, 0, 1, 1])                   # GeoKeyDirectory: RasterPixelIsArea
    off_tag, cnt_tag = (324, 325) if tile else (273, 279)
    add(off_tag, 4, [0] * len(blocks)); add(cnt_tag, 4, [len(b) for b in blocks])
    entries.sort(key=lambda t: t[0])
    ifd_off = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_size
    extra = bytearray()
    placed = {}
    for tag, typ, vals in entries:
        size = struct.calcsize("<" + _TYPE_FMT[typ] * len(vals))
        if size > 4:
            placed[tag] = extra_off + len(extra)
            extra += b"\0" * (size + (size & 1))
    data_off = extra_off + len(extra)
    block_offs, p = [], data_off
    for b in blocks:
        block_offs.append(p)
        p += len(b) + (len(b) & 1)
    out = bytearray(struct.pack("<2sHI", b"II", 42, ifd_off))
    out += struct.pack("<H", len(entries))
    for tag, typ, vals in entries:
        if tag == off_tag:
            vals = tuple(block_offs)
        packed = struct.pack("<" + _TYPE_FMT[typ] * len(vals), *vals)
        if len(packed) > 4:
            o = placed[tag] - extra_off
            extra[o:o + len(packed)] = packed
            out += struct.pack("<HHII", tag, typ, len(vals), placed[tag])
        else:
            out += struct.pack("<HHI", tag, typ, len(vals)) + packed.ljust(4, b"\0")
    out += struct.pack("<I", 0) + bytes(extra)
    for b in blocks:
        out += b + (b"\0" if len(b) & 1 else b"")
    with open(path, "wb") as f:
        f.write(bytes(out))
