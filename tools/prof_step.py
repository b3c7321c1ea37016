"""Run the end-to-end step of bench.py a few times (ncu launch-list / timing target).
usage: python tools/prof_step.py [side] [steps] [cascade]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmerge_b200 import MergeEngine
from deepmerge_b200.synth import CASCADE_TAU, cascade_feats, synth_scene

side = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cascade = len(sys.argv) > 3 and sys.argv[3] == "cascade"
rmul = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # engine sized for rmul x the regions (ids of other row tiles)
dev = torch.device("cuda:0")
sc = synth_scene(side, side, int(100000 * side * side / 1e8), C=4, device=dev)
feats, tau = (cascade_feats(sc), CASCADE_TAU) if cascade else (sc.feats, 0.5)
eng = MergeEngine(side, side, sc.n_regions * rmul, 100, C=4, n_points=sc.feats.shape[0], device=dev)
for _ in range(3):
    r = eng.run(sc.labels, feats, tau, image=sc.image, xs=sc.xs, ys=sc.ys)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    r = eng.run(sc.labels, feats, tau, image=sc.image, xs=sc.xs, ys=sc.ys)
b.record()
torch.cuda.synchronize()
print("ms/step", a.elapsed_time(b) / steps, "rounds", r.rounds, "merges", r.merges)
