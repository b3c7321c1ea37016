"""GPU: the sharded engine (a) with 2, 3 and 4 virtual ranks emulated as THREADS of one process on one GPU (a fake
torch.distributed whose collectives meet on host barriers; no kernel ever waits on another kernel) and (b), when the
box has at least two GPUs, as real torchrun-launched processes over NCCL (tests/sharded_nccl_worker.py) -- compared
with the single-GPU engine and with the oracle."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import oracle_np as o

pytestmark = pytest.mark.gpu


class FakeDist:
    """Minimal torch.distributed look-alike for `world` threads in one process."""

    class ReduceOp:
        SUM, MAX, MIN = "sum", "max", "min"

    def __init__(self, world):
        import torch
        self.world, self.torch = world, torch
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.local = threading.local()

    def get_world_size(self, group=None):
        return self.world

    def get_rank(self, group=None):
        return self.local.rank

    def _exchange(self, t):
        self.torch.cuda.synchronize()
        self.slots[self.local.rank] = t.clone()
        self.bar.wait()
        vals = [s.clone() for s in self.slots]
        self.bar.wait()
        return vals

    def all_gather(self, out_list, t, group=None):
        for dst, src in zip(out_list, self._exchange(t)):
            dst.copy_(src)

    def all_reduce(self, t, op="sum", group=None):
        vals = self._exchange(t)
        acc = vals[0].clone()
        for v in vals[1:]:                      # fixed rank order
            acc = acc + v if op == "sum" else (self.torch.maximum(acc, v) if op == "max" else self.torch.minimum(acc, v))
        t.copy_(acc)


@pytest.mark.parametrize("world,H,W,R,tau", [(2, 260, 384, 500, 0.5), (3, 301, 640, 1500, 0.5), (4, 203, 512, 900, 0.5),
                                              (4, 64, 256, 700, 0.5), (3, 240, 384, 600, 30.0)])
def test_sharded_equals_single_gpu_and_oracle(cuda, world, H, W, R, tau):
    import torch
    from deepmerge_b200 import merge_scene
    from deepmerge_b200.sharded import ShardedMergeEngine, points_in_tile, tile_bounds
    sc = o.synth_scene(H, W, R, C=4)
    n, D = sc["n_regions"], sc["feats"].shape[1]
    want = o.merge_scene(sc["labels"], n, sc["region_of_point"], sc["feats"], tau=tau)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    fd = FakeDist(world)
    results, errors = [None] * world, []

    def rank_main(rank):
        try:
            fd.local.rank = rank
            torch.cuda.set_device(cuda)
            y0, y1 = tile_bounds(H, world, rank)
            last = rank == world - 1
            labels = T(sc["labels"][y0:y1 + (0 if last else 1)])
            image = T(sc["image"][y0:y1])
            mine = points_in_tile(torch.from_numpy(sc["ys"]), y0, y1).numpy()
            eng = ShardedMergeEngine(H, W, n, D, 4, len(mine), fd, cuda)
            res = eng.run(labels, T(sc["feats"][mine]), tau, image_tile=image, xs_local=T(sc["xs"][mine]),
                          ys_local_rel=T(sc["ys"][mine] - y0))
            torch.cuda.synchronize()
            results[rank] = (res.labels.cpu().numpy(), res.root.cpu().numpy(), res.rounds, res.merges,
                             res.area.cpu().numpy(), res.perimeter.cpu().numpy(),
                             res.edge_keys.cpu().numpy().view(np.uint64), eng.eng.bsum.cpu().numpy().view(np.uint64))
        except Exception as e:          # surface thread failures in the test
            errors.append(e)
            fd.bar.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    full = np.concatenate([r[0] for r in results])
    assert np.array_equal(full, want["labels"])                       # final label map bit exact
    roots = np.unique(want["root"])
    s, _ = o.pool_bands(sc["labels"], sc["image"], n)
    for r in results:
        assert np.array_equal(r[1], want["root"]) and r[2] == want["rounds"] and r[3] == want["merges"]
        assert np.array_equal(r[4][roots], want["area"][roots]) and np.array_equal(r[5][roots], want["perim"][roots])
        assert np.array_equal(r[6], want["keys"])
        assert np.array_equal(r[7], s)                                # band sums: integer all-reduce, exact
    single = merge_scene(T(sc["labels"]), T(sc["feats"]), tau, n_regions=n, image=T(sc["image"]), xs=T(sc["xs"]), ys=T(sc["ys"]))
    assert np.array_equal(single.labels.cpu().numpy(), full)


def test_sharded_row_slot_overflow_is_reported(cuda):
    """The sparse row exchange uses fixed-capacity slots; a slot that is too small must raise, never truncate."""
    import torch
    from deepmerge_b200.sharded import ShardedMergeEngine, points_in_tile, tile_bounds
    H, W, R, world = 120, 256, 400, 2
    sc = o.synth_scene(H, W, R, C=4)
    n, D = sc["n_regions"], sc["feats"].shape[1]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    fd = FakeDist(world)
    errors = []

    def rank_main(rank):
        try:
            fd.local.rank = rank
            torch.cuda.set_device(cuda)
            y0, y1 = tile_bounds(H, world, rank)
            last = rank == world - 1
            mine = points_in_tile(torch.from_numpy(sc["ys"]), y0, y1).numpy()
            eng = ShardedMergeEngine(H, W, n, D, 4, len(mine), fd, cuda, row_capacity=2)
            eng.run(T(sc["labels"][y0:y1 + (0 if last else 1)]), T(sc["feats"][mine]), 0.5, image_tile=T(sc["image"][y0:y1]),
                    xs_local=T(sc["xs"][mine]), ys_local_rel=T(sc["ys"][mine] - y0))
        except RuntimeError as e:
            errors.append(str(e))
        except Exception as e:
            errors.append("unexpected: %r" % (e,))
            fd.bar.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert len(errors) == world and all("slot overflow" in e for e in errors), errors


def test_real_nccl_ranks_equal_the_oracle(cuda):
    """One process per GPU over real NCCL (torchrun, 2 ranks; 4 when the box has them): label tiles, root table, rounds,
    merges, merged areas and the gathered edge list equal the oracle's."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    for world in ([2, 4] if n >= 4 else [2]):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(29613 + world),
                            os.path.join(here, "sharded_nccl_worker.py")], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "NCCL PARITY OK" in r.stdout, (world, r.stdout[-2000:], r.stderr[-2000:])
