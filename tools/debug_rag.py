import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle_np as o
from deepmerge_b200 import build_rag
H, W, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(H * 1000 + W + C)
R = 37
small = rng.integers(0, R, size=((H + 5) // 6, (W + 6) // 7)).astype(np.int32)
L = np.kron(small, np.ones((6, 7), np.int32))[:H, :W].copy()
L[rng.random((H, W)) < 0.02] = rng.integers(0, R)
dev = torch.device("cuda:0")
rag = build_rag(torch.from_numpy(L).to(dev), R)
keys, blen, area, per = o.build_rag(L, R)
gk = rag.edge_keys.cpu().numpy().view(np.uint64); gb = rag.boundary_len.cpu().numpy().view(np.uint32)
print("keys equal", np.array_equal(gk, keys), "area equal", np.array_equal(rag.area.cpu().numpy(), area),
      "sum blen", gb.sum(), blen.sum())
if np.array_equal(gk, keys):
    bad = np.nonzero(gb != blen)[0]
    print("n bad", bad.size)
    for i in bad[:10]:
        lo, hi = int(keys[i] >> np.uint64(32)), int(keys[i] & np.uint64(0xffffffff))
        # where are these pairs?
        hh = np.argwhere(((L[:, :-1] == lo) & (L[:, 1:] == hi)) | ((L[:, :-1] == hi) & (L[:, 1:] == lo)))
        vv = np.argwhere(((L[:-1] == lo) & (L[1:] == hi)) | ((L[:-1] == hi) & (L[1:] == lo)))
        print("edge", lo, hi, "gpu", gb[i], "want", blen[i], "h pairs (y,x):", hh[:6].tolist(), "v pairs:", vv[:6].tolist())
