"""Device-side generator of the synthetic scenes of SURVEY.md section 8(d) (bench / test
utility).  Bit-identical to oracle/oracle_np.py's generator; the GPU tests check that."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ._lib import synth_lib
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
    L = synth_lib()
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


CASCADE_TAU = 0.5
CASCADE_AMPL = (0.25, 0.27386127, 0.27386127, 8.0)


def cascade_feats(scene: Scene, D=100):
    """Embeddings of the multi-round workload (bench.py `multi_round`; the same construction as the oracle's
    synth_cascade_feats, bit for bit): four one-hot parts per point -- region, object, group of 4 x 4 objects, top --
    whose amplitudes make regions merge into objects in round 1, objects into groups in round 2 (the merged means have
    lost most of the region part), groups into tops in round 3, with tau = CASCADE_TAU."""
    if D < 100:
        raise ValueError("the cascade construction uses 100 embedding dimensions")
    dev = scene.labels.device
    g = grid_pitch(scene.H, scene.W, scene.n_regions)
    ncx, ncx_o = -(-scene.W // g), -(-scene.W // (4 * g))
    rop = scene.region_of_point.to(torch.int64)
    ok = rop >= 0
    r = torch.where(ok, rop, torch.zeros_like(rop))
    cx, cy = r % ncx, r // ncx
    o = scene.region_obj.to(torch.int64)[r]
    ox, oy = o % ncx_o, o // ncx_o
    sx, sy = ox // 4, oy // 4
    tx, ty = sx // 4, sy // 4
    idx = [(cx % 8) + 8 * (cy % 5), 40 + (ox % 6) + 6 * (oy % 5), 70 + (sx % 5) + 5 * (sy % 4), 90 + (tx % 5) + 5 * (ty % 2)]
    f = torch.zeros((rop.shape[0], D), dtype=torch.float32, device=dev)
    rows = torch.arange(rop.shape[0], device=dev)
    for k, a in zip(idx, CASCADE_AMPL):
        f[rows, k] = a
    f[~ok] = 0
    return f
