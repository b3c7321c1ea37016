"""Worker of tests/test_gpu_sharded.py::test_real_nccl_ranks_equal_the_oracle (TEST INFRASTRUCTURE): launched by
torchrun with one process per GPU.  Every rank runs ShardedMergeEngine over real torch.distributed / NCCL on its row
tile of an oracle-generated scene; rank 0 gathers the tiles and compares with the CPU oracle and with the single-GPU
engine, bit for bit.  Exit code 0 and the line "NCCL PARITY OK" on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_np as o                                     # noqa: E402  (the checker)
from deepmerge_b200 import merge_scene                                # noqa: E402
from deepmerge_b200.sharded import ShardedMergeEngine, points_in_tile, position_checksum, tile_bounds   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # the last scene is the multi-round workload (three rounds: the captured round body is replayed, its exchange buffers
    # are reused round after round); every scene is run twice by the same engine (graphs captured in the first run)
    for H, W, R, tau in ((260, 384, 500, 0.5), (1203, 1024, 5000, 0.5), (240, 384, 600, 30.0), (600, 800, 4000, "cascade")):
        sc = o.synth_scene(H, W, R, C=4)
        if tau == "cascade":
            from deepmerge_b200.synth import CASCADE_TAU
            sc["feats"] = o.synth_cascade_feats(sc["region_of_point"], sc["region_obj"], H, W, R)
            tau = CASCADE_TAU
        n, D = sc["n_regions"], sc["feats"].shape[1]
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        y0, y1 = tile_bounds(H, world, rank)
        last = rank == world - 1
        mine = points_in_tile(torch.from_numpy(sc["ys"]), y0, y1).numpy()
        eng = ShardedMergeEngine(H, W, n, D, 4, len(mine), dist, dev)
        tile = (T(sc["labels"][y0:y1 + (0 if last else 1)]), T(sc["feats"][mine]))
        kw = dict(image_tile=T(sc["image"][y0:y1]), xs_local=T(sc["xs"][mine]), ys_local_rel=T(sc["ys"][mine] - y0))
        eng.run(*tile, tau, gather_outputs=False, **kw)
        res = eng.run(*tile, tau, gather_outputs=True, **kw)
        cs = torch.tensor([position_checksum(res.labels, y0 * W), position_checksum(res.root), res.rounds, res.merges],
                          dtype=torch.int64, device=dev)
        every = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(every, cs)
        if rank == 0:
            want = o.merge_scene(sc["labels"], n, sc["region_of_point"], sc["feats"], tau=tau)
            if (H, W, R) == (600, 800, 4000):
                assert want["rounds"] == 3, want["rounds"]
            single = merge_scene(T(sc["labels"]), T(sc["feats"]), tau, n_regions=n, image=T(sc["image"]), xs=T(sc["xs"]),
                                 ys=T(sc["ys"]))
            assert np.array_equal(single.labels.cpu().numpy(), want["labels"])
            wl, wr = T(want["labels"]), T(want["root"])
            for r in range(world):
                a, b = tile_bounds(H, world, r)
                exp = [position_checksum(wl[a:b], a * W), position_checksum(wr), want["rounds"], want["merges"]]
                if [int(x) for x in every[r].tolist()] != exp:
                    ok = False
                    print("MISMATCH scene", (H, W, R, tau), "rank", r, every[r].tolist(), exp, flush=True)
            roots = np.unique(want["root"])
            ok = ok and np.array_equal(res.area.cpu().numpy()[roots], want["area"][roots])
            ok = ok and np.array_equal(res.edge_keys.cpu().numpy().view(np.uint64), want["keys"])
    # the pair-MLP as the scorer of the distributed loop (a hand-made "same object?" network with wide logit margins)
    from deepmerge_b200 import PackedMLP
    H, W, R = 400, 512, 900
    sc = o.synth_scene(H, W, R, C=4)
    n, D = sc["n_regions"], sc["feats"].shape[1]
    W1 = np.zeros((2 * D, 2 * D), np.float32)
    for d in range(D):
        W1[d, d], W1[d, D + d] = 1, -1
        W1[D + d, d], W1[D + d, D + d] = -1, 1
    W3 = np.zeros((2, 2 * D), np.float32)
    W3[0] = 1.0
    Wm = [W1, np.zeros(2 * D, np.float32), np.eye(2 * D, dtype=np.float32), np.zeros(2 * D, np.float32), W3,
          np.array([0.0, 4.0], np.float32)]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    y0, y1 = tile_bounds(H, world, rank)
    last = rank == world - 1
    mine = points_in_tile(torch.from_numpy(sc["ys"]), y0, y1).numpy()
    eng = ShardedMergeEngine(H, W, n, D, 4, len(mine), dist, dev)
    res = eng.run(T(sc["labels"][y0:y1 + (0 if last else 1)]), T(sc["feats"][mine]), 0.0, image_tile=T(sc["image"][y0:y1]),
                  xs_local=T(sc["xs"][mine]), ys_local_rel=T(sc["ys"][mine] - y0), gather_outputs=False,
                  mlp=PackedMLP(*[T(w) for w in Wm]))
    want = o.merge_scene(sc["labels"], n, sc["region_of_point"], sc["feats"], mlp=tuple(Wm))
    mine_ok = bool(np.array_equal(res.labels.cpu().numpy(), want["labels"][y0:y1])) and \
        bool(np.array_equal(res.root.cpu().numpy(), want["root"])) and res.merges == want["merges"] and want["merges"] > 0
    okt = torch.tensor([1 if mine_ok else 0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0 and not int(okt.item()):
        print("MISMATCH in the MLP-scored sharded merge", flush=True)
    ok = ok and bool(int(okt.item()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("NCCL PARITY OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
