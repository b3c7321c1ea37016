"""CPU, gloo, world_size 2: the host logic of the row-tile sharding (tile bounds, variable
length edge-list all-gather, statistics all-reduce) composed with the oracle's tile RAG must
reproduce the whole-scene graph."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepmerge_b200.sharded import allreduce_sum_, gather_edge_lists, points_in_tile, tile_bounds
from oracle import oracle_np as o


def test_tile_bounds_cover_every_row_once():
    for H in (1, 7, 100, 10001):
        for G in (1, 2, 3, 8):
            rows = []
            for g in range(G):
                y0, y1 = tile_bounds(H, G, g)
                rows += list(range(y0, y1))
            assert rows == list(range(H))
            sizes = [tile_bounds(H, G, g)[1] - tile_bounds(H, G, g)[0] for g in range(G)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, R, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = o.synth_scene(H, W, R, C=3)
    L, n = sc["labels"], sc["n_regions"]
    y0, y1 = tile_bounds(H, world, rank)
    last = rank == world - 1
    tile = L[y0:y1 + (0 if last else 1)]
    keys, blen, area, per = o.build_rag(tile, n, top_border=rank == 0, bottom_border=last, own_rows=y1 - y0)
    # the product's perimeter = border + incident boundary lengths; here: exchange the oracle's tile perimeters
    k, b = gather_edge_lists(torch.from_numpy(keys.view(np.int64)), torch.from_numpy(blen.astype(np.int64)), dist)
    stats = torch.from_numpy(np.concatenate([area, per]))
    allreduce_sum_(stats, dist)
    s, q = o.pool_bands(tile[: y1 - y0], sc["image"][y0:y1], n)
    bands = torch.from_numpy(np.concatenate([s.ravel(), q.ravel()]).astype(np.int64))
    allreduce_sum_(bands, dist)
    # points go with the tile that contains them; partial pooled sums are all-reduced
    mine = points_in_tile(torch.from_numpy(sc["ys"]), y0, y1).numpy()
    rop = tile[sc["ys"][mine] - y0, sc["xs"][mine]]
    off, ids = o.csr_from_region_of_point(rop, n)
    ps, pc, _ = o.pool_points_csr(off, ids, sc["feats"][mine])
    ps_t, pc_t = torch.from_numpy(ps), torch.from_numpy(pc.astype(np.int64))
    allreduce_sum_(ps_t, dist)
    allreduce_sum_(pc_t, dist)
    if rank == 0:
        uk, inv = np.unique(k.numpy().view(np.uint64), return_inverse=True)
        ub = np.bincount(inv, weights=b.numpy().astype(np.float64)).astype(np.int64)
        np.savez(out, keys=uk, blen=ub, stats=stats.numpy(), bands=bands.numpy(), psum=ps_t.numpy(), pcnt=pc_t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W,R", [(121, 96, 60), (64, 200, 300)])
def test_two_rank_gloo_exchange_reproduces_whole_scene(tmp_path, H, W, R):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), H, W, R, out), nprocs=2, join=True)
    got = np.load(out)
    sc = o.synth_scene(H, W, R, C=3)
    n = sc["n_regions"]
    keys, blen, area, per = o.build_rag(sc["labels"], n)
    assert np.array_equal(got["keys"], keys) and np.array_equal(got["blen"], blen)
    assert np.array_equal(got["stats"], np.concatenate([area, per]))
    s, q = o.pool_bands(sc["labels"], sc["image"], n)
    assert np.array_equal(got["bands"], np.concatenate([s.ravel(), q.ravel()]).astype(np.int64))
    off, ids = o.csr_from_region_of_point(sc["region_of_point"], n)
    ps, pc, _ = o.pool_points_csr(off, ids, sc["feats"])
    assert np.array_equal(got["pcnt"], pc)
    np.testing.assert_allclose(got["psum"], ps, rtol=1e-5, atol=1e-5)       # fp32 sums: order differs across tiles


def _slots_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepmerge_b200.sharded import all_gather_slots
    # fixed-capacity slot = [count | payload]: what the sharded engine ships instead of variable-length lists
    slot = torch.full((6,), -1, dtype=torch.int64)
    n = 2 + rank
    slot[0] = n
    slot[1:1 + n] = torch.arange(n) + 100 * rank
    g = all_gather_slots(slot, dist).view(world, 6)
    parent = torch.tensor([0, 1, 2, 3, 4, 5], dtype=torch.int32)
    if rank == 0:
        parent[3] = 1                    # rank 0 learnt 3 ~ 1
    else:
        parent[3] = 2                    # rank 1 learnt 3 ~ 2: the min keeps 1, rank 1 re-unions 2 ~ 3 next iteration
    dist.all_reduce(parent, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(out, g=g.numpy(), parent=parent.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_slot_gather_and_parent_min(tmp_path):
    out = str(tmp_path / "s.npz")
    mp.spawn(_slots_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert got["g"][0].tolist() == [2, 0, 1, -1, -1, -1] and got["g"][1].tolist() == [3, 100, 101, 102, -1, -1]
    assert got["parent"].tolist() == [0, 1, 2, 1, 4, 5]
