"""Times dm_region_bbox on the configs[1] label raster (N2, not part of the merge step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepmerge_b200 import raster as rs, synth

sc = synth.synth_scene(10000, 10000, 100000, C=0, with_image=False)
lab, R = sc.labels, sc.n_regions
for _ in range(3):
    rs.region_bbox(lab, R)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(10):
    rs.region_bbox(lab, R)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"region_bbox 10k x 10k R={R}: {ms:.3f} ms  ({lab.numel() * 4 / ms / 1e6:.0f} GB/s of label reads)")
