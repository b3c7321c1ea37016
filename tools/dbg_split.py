"""Debug helper: one dm_rag_build through the warp-specialised kernel on a long-strip scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DM_RAG_KERNEL"] = "split"
import numpy as np, torch
from deepmerge_b200 import build_rag
from deepmerge_b200.synth import synth_scene
H, W, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sc = synth_scene(H, W, R, C=4, device=torch.device("cuda:0"))
rag = build_rag(sc.labels, sc.n_regions, sc.image)
torch.cuda.synchronize()
print("edges", rag.n_edges, "area sum", int(rag.area.sum()), "expect", H * W)
