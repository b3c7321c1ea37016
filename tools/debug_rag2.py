import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle_np as o
from deepmerge_b200 import build_rag
dev = torch.device("cuda:0")
def run(L, R=40, tag=""):
    rag = build_rag(torch.from_numpy(L).to(dev), R)
    keys, blen, area, per = o.build_rag(L, R)
    gk = rag.edge_keys.cpu().numpy().view(np.uint64); gb = rag.boundary_len.cpu().numpy().view(np.uint32)
    ok = np.array_equal(gk, keys) and np.array_equal(gb, blen)
    print(tag, "OK" if ok else f"BAD gpu={dict(zip(gk.tolist(), gb.tolist()))} want={dict(zip(keys.tolist(), blen.tolist()))}")
H, W = 40, 256
for (y, x) in [(5, 9), (5, 10), (5, 8), (5, 11), (7, 10), (8, 10), (15, 10), (16, 10), (5, 127), (5, 128), (5, 124)]:
    L = np.ones((H, W), np.int32); L[y, x] = 8
    run(L, tag=f"speckle({y},{x})")
L = np.ones((H, W), np.int32); L[5, 9] = 8; L[5, 50] = 8; run(L, tag="two speckles same row")
L = np.ones((H, W), np.int32); L[5, 9] = 8; L[20, 9] = 8; run(L, tag="two speckles same col")
L = np.ones((H, W), np.int32); L[5, 9] = 8; L[7, 9] = 8; run(L, tag="two speckles col gap1")
L = np.ones((H, W), np.int32); L[5, 9] = 8; L[6, 9] = 9; run(L, tag="speckles 8 over 9")
L = np.ones((H, W), np.int32); L[:, 100:] = 2; L[5, 9] = 8; run(L, tag="two regions + speckle")
L = np.ones((H, W), np.int32); L[:, 100:] = 2; L[20:, :] = 3; run(L, tag="three regions")
