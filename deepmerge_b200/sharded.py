"""Row-tile sharding of one scene across the GPUs of a box.

Rank g owns rows [g*H/G, (g+1)*H/G) of labels / image plus ONE halo row below: the vertical
pixel pair (y, y+1) belongs to the tile that owns y, so nothing is counted twice.  Region ids
are global.  One exchange step per phase, with torch.distributed as plumbing (NCCL on the
GPUs; the same code runs under gloo on CPU tensors, which is how the host logic is tested):

  1. per-tile unique (key, boundary_len) lists: all_gather_into_tensor of fixed-capacity slots plus
     the device-side counts -> dm_edges_concat + dm_edges_sort_unique merge them into the global
     edge list on every rank (replicated); no length ever travels to the host;
  2. per-region partial statistics: all_reduce(sum) of [area | border | band sums | band
     sums of squares] (int64) and of the pooled embedding sums / point counts;
  3. graph-level work (scoring, union-find merge loop) is small and runs replicated and
     deterministic on every rank -- no further collectives;
  4. the final relabel is tile-local with the replicated root LUT.

The pooled fp32 sums of a region that spans tiles are reduced in NCCL's order, not in the
reference's point order, so they can differ from the single-GPU result in the last bits;
integer outputs are identical and merge decisions agree outside the stated guard band.
"""
from __future__ import annotations

import json
import os
import time
from typing import Optional

import torch


def tile_bounds(H: int, world: int, rank: int):
    """Rows [y0, y1) owned by `rank`; tiles differ by at most one row."""
    base, extra = divmod(H, world)
    y0 = rank * base + min(rank, extra)
    y1 = y0 + base + (1 if rank < extra else 0)
    return y0, y1


def gather_edge_lists(keys: torch.Tensor, lens: torch.Tensor, dist, group=None):
    """all_gather of variable-length per-tile edge lists -> (keys_cat, lens_cat) with the valid
    entries of every rank in rank order.  Lengths are exchanged first, payloads are padded."""
    world = dist.get_world_size(group)
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pk = torch.zeros(m, dtype=keys.dtype, device=keys.device)
    pl = torch.zeros(m, dtype=lens.dtype, device=lens.device)
    pk[: keys.shape[0]] = keys
    pl[: lens.shape[0]] = lens
    gk = [torch.empty_like(pk) for _ in range(world)]
    gl = [torch.empty_like(pl) for _ in range(world)]
    dist.all_gather(gk, pk, group=group)
    dist.all_gather(gl, pl, group=group)
    return torch.cat([g[:c] for g, c in zip(gk, counts)]), torch.cat([g[:c] for g, c in zip(gl, counts)])


def allreduce_sum_(t: torch.Tensor, dist, group=None):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def points_in_tile(ys: torch.Tensor, y0: int, y1: int):
    """Indices of the sample points whose row falls in [y0, y1): they go with that tile."""
    return torch.nonzero((ys >= y0) & (ys < y1)).flatten()


def all_gather_slots(t: torch.Tensor, dist, group=None):
    """all_gather of equal-sized slots into one tensor [world * len(t)] (one NCCL call, no host sync)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * t.shape[0], dtype=t.dtype, device=t.device)
    if hasattr(dist, "all_gather_into_tensor"):
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:                                             # minimal stand-ins used by tests
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out.copy_(torch.cat(parts))
    return out


class ShardedMergeEngine:
    """One rank's share of a scene sharded by rows.  Wraps a MergeEngine sized for the tile.

    A step makes no host round trip before the merge loop's own read-back: the per-tile edge lists
    travel as fixed-capacity slots together with their device-side counts, and overflow / bad-label
    flags of the tile pass are checked with that first read-back."""

    def __init__(self, H, W, n_regions, D, C, n_points_local, dist, device, group=None, slot_capacity=None):
        from .raster import MergeEngine, default_edge_capacity
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.H, self.W = H, W
        self.y0, self.y1 = tile_bounds(H, self.world, self.rank)
        self.rows_own = self.y1 - self.y0
        self.has_halo = self.rank < self.world - 1
        # capacity: the global edge list (and the raw per-tile entries) must fit
        cap = default_edge_capacity(n_regions, H, W)
        self.slot_cap = int(slot_capacity) if slot_capacity else cap // self.world + 4 * W + 4096
        self.eng = MergeEngine(self.rows_own, W, n_regions, D, C=C, n_points=n_points_local, edge_capacity=cap,
                               device=device)
        dev = self.eng.dev
        self.tile_counts = torch.zeros(4, dtype=torch.int64, device=dev)
        self.cat_counts = torch.zeros(2, dtype=torch.int64, device=dev)

    def run(self, labels_tile, feats_local, tau, *, image_tile=None, xs_local=None, ys_local_rel=None, max_rounds=64):
        """labels_tile: int32 [rows_own (+1 halo), W]; ys_local_rel are rows relative to the tile."""
        from .raster import MergeResult, _p, _stream
        e, L, dist = self.eng, self.eng.L, self.dist
        with torch.cuda.device(e.dev):
            s = _stream()
            e._rag(labels_tile, image_tile, self.rows_own, self.rank == 0, self.rank == self.world - 1)
            self.tile_counts.copy_(e.counts[:4])
            # (1) edges: gather the per-tile lists (fixed slots + device counts), merge into the replicated global list
            if self.slot_cap > e.cap:
                raise ValueError("slot capacity exceeds the engine's edge capacity")
            gk = all_gather_slots(e.keys[: self.slot_cap], dist, self.group)
            gl = all_gather_slots(e.blen[: self.slot_cap], dist, self.group)
            gc = all_gather_slots(e.counts[0:1], dist, self.group)
            L.check(L.dm_edges_concat(_p(gk), _p(gl), _p(gc), self.world, self.slot_cap, _p(e.keys), _p(e.blen), e.cap,
                                      _p(self.cat_counts), s), "dm_edges_concat")
            e.counts[:4].zero_()
            L.check(L.dm_edges_sort_unique(_p(e.keys), _p(e.blen), _p(self.cat_counts), e.cap, e.R, _p(e.counts), _p(e.ws),
                                           e.ws_bytes, s), "dm_edges_sort_unique")
            # tile-pass flags ride along with the merge loop's first read-back (counts[2] overflow, counts[3] bad label)
            e.counts[2:3].copy_(torch.maximum(self.tile_counts[2:3], self.cat_counts[1:2]))
            e.counts[3:4].copy_(self.tile_counts[3:4])
            e.counts[1:2].copy_(self.tile_counts[1:2])
            # (2) region statistics: one all-reduce over the fused int64 buffer, then the perimeter
            allreduce_sum_(e.stats, dist, self.group)
            L.check(L.dm_perimeter(_p(e.keys), _p(e.blen), _p(e.counts), e.cap, _p(e.border), _p(e.perim), e.R, s),
                    "dm_perimeter")
            # (3) embeddings of the tile's own points -> partial sums -> all-reduce
            e._pool(labels_tile, xs_local, ys_local_rel, None, feats_local)
            allreduce_sum_(e.sum, dist, self.group)
            allreduce_sum_(e.cnt, dist, self.group)
            # (4) replicated merge loop, (5) tile-local relabel
            try:
                rounds, merges = e._merge_loop(tau, max_rounds)
            except OverflowError as ov:
                raise RuntimeError("tile edge list overflow (capacity %d / slot %d, needed %s)" % (e.cap, self.slot_cap, ov.args[0]))
            L.check(L.dm_relabel(_p(labels_tile), self.rows_own, self.W, labels_tile.stride(0), _p(e.parent), e.R, _p(e.out),
                                 self.W, s), "dm_relabel")
            Ef = int(e.host_counts[0])
        return MergeResult(e.out[: self.rows_own], e.parent, rounds, merges, e.keys[:Ef], e.blen[:Ef], e.scores[:Ef], e.area,
                           e.perim, e.sum, e.cnt)


# ----------------------------------------------------------------------------------------------
# bench leg for N > 1 (launched by torchrun, one rank per GPU)
# ----------------------------------------------------------------------------------------------
SHARDED_CFG = dict(H=40000, W=40000, R=1000000, C=4, P=4, D=100, tau=0.5, seed=1234)
SHARDED_WORKLOAD = "configs[2]: 40k x 40k scene, ~1M segments, row-tile sharded across N B200 with NCCL boundary exchange"


def bench_sharded(args, cfg1, workload1, dist, dev, ClockSampler, measured_peaks):
    """Weak scaling: every rank owns a tile of H_1 x W_1 / ... -- the N-GPU scene is the single-GPU
    workload's height times N (same tile per GPU), so per-GPU work is fixed as N grows."""
    import numpy as np
    from . import _lib
    from .synth import grid_pitch, synth_scene
    from .raster import _p, _stream, points_region

    world, rank = dist.get_world_size(), dist.get_rank()
    L = _lib.lib()
    cfg = dict(cfg1)
    if args.side:
        cfg.update(H=args.side, W=args.side, R=max(4, int(round(cfg1["R"] * args.side * args.side / (cfg1["H"] * cfg1["W"])))))
    # weak scaling: N tiles of the single-GPU scene stacked vertically
    H, W, C, D, P = cfg["H"] * world, cfg["W"], cfg["C"], cfg["D"], cfg["P"]
    R_target = cfg["R"] * world
    fixed_scene = bool(getattr(args, "config2", False))
    if fixed_scene:                       # BASELINE.json configs[2]: ONE 40k x 40k scene split over the ranks (strong scaling)
        cfg = dict(SHARDED_CFG)
        H, W, C, D, P, R_target = cfg["H"], cfg["W"], cfg["C"], cfg["D"], cfg["P"], cfg["R"]
    y0, y1 = tile_bounds(H, world, rank)
    has_halo = rank < world - 1
    sc = synth_scene(H, W, R_target, C=C, P=P, D=D, seed=cfg["seed"], device=dev, rows=(y0, y1 + (1 if has_halo else 0)))
    R = sc.n_regions
    mine = points_in_tile(sc.ys, y0, y1)
    xs, ys_rel = sc.xs[mine].contiguous(), (sc.ys[mine] - y0).contiguous()
    # embeddings of the tile's points (region_of_point from the tile's own labels)
    rop = points_region(sc.labels[: y1 - y0], xs, ys_rel)
    feats = torch.empty((xs.shape[0], D), dtype=torch.float32, device=dev)
    # feats are hashed per GLOBAL point id so that the scene does not depend on the sharding
    L.check(L.dm_synth_feats(_p(feats), _p(rop), _p(sc.region_obj), _p(mine.contiguous()), xs.shape[0], D, cfg["seed"],
                             _stream()), "dm_synth_feats")
    image = sc.image[: y1 - y0] if sc.image is not None else None
    eng = ShardedMergeEngine(H, W, R, D, C, xs.shape[0], dist, dev)

    def step():
        return eng.run(sc.labels, feats, cfg["tau"], image_tile=image, xs_local=xs, ys_local_rel=ys_rel)

    for _ in range(max(args.warmup, 3)):
        res = step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.dm_launch_count()
    with ClockSampler(dev.index) as clocks:
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            res = step()
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
    ms_local = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = L.dm_launch_count() - launches0

    # end to end with HOST buffers on every rank (tile H2D, label-map tile D2H inside the timed region)
    host = {"labels": sc.labels.cpu().pin_memory(), "image": image.cpu().pin_memory(), "feats": feats.cpu().pin_memory(),
            "xs": xs.cpu().pin_memory(), "ys": ys_rel.cpu().pin_memory()}
    out_host = torch.empty((y1 - y0, W), dtype=torch.int32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        r = eng.run(d["labels"], d["feats"], cfg["tau"], image_tile=d["image"], xs_local=d["xs"], ys_local_rel=d["ys"])
        out_host.copy_(r.labels, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 3)):
        e2e_step()
    dist.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, args.steps // 3)
    t = torch.tensor([e2e_ms, float(h2d), float(out_host.numel() * 4)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, kind = measured_peaks()
        n_roots = int((res.root == torch.arange(R, device=dev, dtype=torch.int32)).sum())
        line = {
            "metric": "megapixels/sec end-to-end (RAG+pool+score+merge+relabel)", "value": H * W / ms / 1e3, "unit": "Mpx/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if fixed_scene else "weak", "vs_baseline": None, "dtype": "int32/u8 index + fp32 scores",
            "data": "synthetic",
            "config": {"workload": (SHARDED_WORKLOAD if fixed_scene else
                                    f"{workload1} -- one such tile per GPU, stacked by rows ({H}x{W}), row-tile sharded with "
                                    "NCCL edge-list all-gather + region-statistics all-reduce"),
                       "H": H, "W": W, "bands": C, "segments": R,
                       "points": int(sc.xs.shape[0]), "embed_dim": D, "tau": cfg["tau"], "parallelism": f"row-tiles x{world}",
                       "l2_policy": "inputs larger than L2, no flush needed"},
            "merged_edges_per_s": res.merges / (ms * 1e-3), "segments_after": n_roots, "rounds": res.rounds,
            "e2e": {"value": H * W / float(tmax[0]) / 1e3, "unit": "Mpx/s", "ms_per_step": float(tmax[0]),
                    "h2d_bytes_per_step": int(t[1]), "d2h_bytes_per_step": int(t[2])},
            "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None, "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
