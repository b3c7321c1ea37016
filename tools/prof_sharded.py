"""Section timing of the distributed merge loop (torchrun, N ranks): wraps torch.distributed collectives and the
library calls of one ShardedMergeEngine.run with CUDA events.  Debug tool."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from deepmerge_b200 import _lib
from deepmerge_b200.sharded import ShardedMergeEngine, tile_bounds, points_in_tile
from deepmerge_b200.synth import synth_scene
from deepmerge_b200.raster import _p, _stream, points_region

local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
L = _lib.lib()
H, W, R_t = 10000 * world, 10000, 100000 * world
y0, y1 = tile_bounds(H, world, rank); halo = rank < world - 1
sc = synth_scene(H, W, R_t, C=4, device=dev, rows=(y0, y1 + (1 if halo else 0)))
mine = points_in_tile(sc.ys, y0, y1); xs, ys = sc.xs[mine].contiguous(), (sc.ys[mine] - y0).contiguous()
rop = points_region(sc.labels[: y1 - y0], xs, ys)
feats = torch.empty((xs.shape[0], 100), device=dev)
SL = _lib.synth_lib()
SL.check(SL.dm_synth_feats(_p(feats), _p(rop), _p(sc.region_obj), _p(mine.contiguous()), xs.shape[0], 100, 1234, _stream()), "f")
eng = ShardedMergeEngine(H, W, sc.n_regions, 100, 4, xs.shape[0], dist, dev)
run = lambda: eng.run(sc.labels, feats, 0.5, image_tile=sc.image[: y1 - y0], xs_local=xs, ys_local_rel=ys)
for _ in range(3): run()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5): run()
torch.cuda.synchronize()
print(rank, "ms/step", (time.perf_counter() - t0) * 200)
# wall-clock sections via monkeypatching
acc = collections.Counter(); cnt = collections.Counter()
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
        acc[label] += time.perf_counter() - t; cnt[label] += 1; return r
    setattr(obj, name, g)
for n in ("all_reduce", "all_gather_into_tensor"): wrap(dist, n, "nccl:" + n)
for n in [k for k in L.protos if k.startswith("dm_")]:
    wrap(L, n, n)
wrap(eng.eng, "_rag", "_rag(total)"); wrap(eng.eng, "_pool", "_pool(total)")
NREP = 5
run(); acc.clear(); cnt.clear()
t0 = time.perf_counter()
for _ in range(NREP): run()
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / NREP
if rank == 0:
    print("serialised total ms", tot * 1e3, "(sections: mean of %d runs, a synchronisation on both sides of every call)" % NREP)
    for k, v in acc.most_common(40): print(f"{v*1e3/NREP:8.3f} ms x{cnt[k]//NREP:3d} {k}")
dist.barrier(); dist.destroy_process_group()
