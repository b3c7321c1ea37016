"""Run only the pair-MLP scorer (tcgen05) on the config-2 edge list a few times (ncu / timing target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmerge_b200 import PackedMLP, build_rag, score_mlp
from deepmerge_b200.synth import synth_scene

side = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
sc = synth_scene(side, side, int(100000 * side * side / 1e8), C=0, device=dev, with_image=False)
rag = build_rag(sc.labels, sc.n_regions)
D, hid, n_out = 100, 250, 2
g = torch.Generator().manual_seed(1)
mk = lambda *sh: (torch.randn(*sh, generator=g) / sh[-1] ** 0.5).to(dev)
mlp = PackedMLP(mk(hid, 2 * D), mk(hid), mk(hid, hid), mk(hid), mk(n_out, hid), mk(n_out))
mean = torch.randn(sc.n_regions, D, device=dev)
E = rag.n_edges
out = torch.empty((E, n_out), device=dev)
n = torch.tensor([E], dtype=torch.int64, device=dev)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record()
    score_mlp(mean, rag.edge_keys, mlp, n_edges_dev=n, out=out)
    b.record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
fl = 2.0 * E * (2 * D * hid + hid * hid + hid * n_out)
print("edges", E, "ms", ms, "useful TFLOP/s", [fl / m / 1e9 for m in ms])
