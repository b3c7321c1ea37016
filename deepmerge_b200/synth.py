"""Device-side generator of the synthetic scenes of SURVEY.md section 8(d) (bench / test
utility).  Bit-identical to oracle/oracle_np.py's generator; the GPU tests check that."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ._lib import lib
from .raster import _p, _stream, points_region


def grid_pitch(H, W, R):
    return max(1, int(round((H * W / max(R, 1)) ** 0.5)))


@dataclass
class Scene:
    labels: torch.Tensor           # int32 [rows, W]
    image: torch.Tensor            # uint8 [rows, W, C]
    xs: torch.Tensor               # int32 [N]
    ys: torch.Tensor
    region_of_point: torch.Tensor  # int32 [N]
    feats: torch.Tensor            # float32 [N, D]
    region_obj: torch.Tensor       # int32 [R]
    n_regions: int
    H: int
    W: int


def synth_scene(H, W, R, C=4, P=4, D=100, seed=1234, device=None, rows=None, with_image=True) -> Scene:
    """Whole scene (rows=None) or the row range rows=(y0, y1) of it.  Points / feats are always
    the whole scene's (they are small); region_of_point needs the whole label raster, so for a
    row range it is left empty."""
    L = lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = grid_pitch(H, W, R)
    ncx, ncy = -(-W // g), -(-H // g)
    nreg = ncx * ncy
    y0, y1 = (0, H) if rows is None else rows
    with torch.cuda.device(dev):
        s = _stream()
        labels = torch.empty((y1 - y0, W), dtype=torch.int32, device=dev)
        L.check(L.dm_synth_labels(_p(labels), y0, y1 - y0, H, W, W, g, seed, s), "dm_synth_labels")
        robj = torch.empty(nreg, dtype=torch.int32, device=dev)
        L.check(L.dm_synth_region_objects(_p(robj), H, W, g, seed, s), "dm_synth_region_objects")
        image = None
        if with_image and C > 0:
            image = torch.empty((y1 - y0, W, C), dtype=torch.uint8, device=dev)
            L.check(L.dm_synth_image(_p(image), _p(labels), y0, y1 - y0, W, W, C, _p(robj), seed, s), "dm_synth_image")
        xs = torch.empty(nreg * P, dtype=torch.int32, device=dev)
        ys = torch.empty(nreg * P, dtype=torch.int32, device=dev)
        L.check(L.dm_synth_points(_p(xs), _p(ys), H, W, g, P, seed, s), "dm_synth_points")
        rop = feats = None
        if rows is None:
            rop = points_region(labels, xs, ys)
            feats = torch.empty((nreg * P, D), dtype=torch.float32, device=dev)
            L.check(L.dm_synth_feats(_p(feats), _p(rop), _p(robj), None, nreg * P, D, seed, s), "dm_synth_feats")
    return Scene(labels, image, xs, ys, rop, feats, robj, nreg, H, W)
