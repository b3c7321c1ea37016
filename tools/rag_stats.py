"""Eviction / path statistics of the RAG raster kernel (needs the -DDM_RAG_STATS build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmerge_b200 import _lib
from deepmerge_b200.raster import _p, _stream
from deepmerge_b200.synth import synth_scene

side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else int(100000 * side * side / 1e8)
C = 4
_lib._LIB = _lib.Library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stats", "libdm_stats.so"))
L = _lib.lib()
dev = torch.device("cuda:0")
sc = synth_scene(side, side, R, C=C, device=dev)
n = sc.n_regions
cap = 8 * n + 100000
area = torch.zeros(n, dtype=torch.int64, device=dev); border = torch.zeros_like(area)
bs = torch.zeros((n, C), dtype=torch.int64, device=dev); bq = torch.zeros_like(bs)
counts = torch.zeros(32 + 2 * 148 * 16, dtype=torch.int64, device=dev)
wsb = L.dm_rag_workspace_bytes(cap); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
L.check(L.dm_rag_scan(_p(sc.labels), side, side, side, side, _p(sc.image), C, side * C, n, 1, 1, _p(area), _p(border), _p(bs),
                      _p(bq), cap, _p(counts), _p(ws), wsb, _stream()), "scan")
c = counts.tolist()
rows = side * ((side + 127) // 128)
names = ["own-miss evict", "uncached add", "prefetch evict", "fast lane-rows", "slow lane-rows", "border lane-rows", "drains", "prefetch: right", "prefetch: 2nd try", "prefetch of last-evicted", "own-miss of last-evicted"]
print("raw", c[1], "warp-rows", rows, "lane-rows", rows * 32)
for nm, v in zip(names, c[8:19]):
    print(f"{nm:18s} {v:12d}  per warp-row {v / rows:8.3f}")

import numpy as np
t = np.array(c[32:]).reshape(-1, 2)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
dur = en - st
print(f"warps {len(t)}  start skew max {st.max():.1f} us  end min/avg/max {en.min():.1f}/{en.mean():.1f}/{en.max():.1f} us  dur min/avg/max {dur.min():.1f}/{dur.mean():.1f}/{dur.max():.1f} us")
cta = dur.reshape(-1, 16) if len(dur) % 16 == 0 else None
if cta is not None:
    m = cta.mean(1); print("per-CTA mean dur: min %.1f avg %.1f max %.1f" % (m.min(), m.mean(), m.max()))
    worst = np.argsort(-m)[:5]; print("slowest CTAs", worst.tolist(), m[worst].round(1).tolist())
    print("within-CTA spread (max-min) avg %.1f" % (cta.max(1) - cta.min(1)).mean())
order = np.argsort(-dur)[:8]; print("slowest warps", order.tolist(), dur[order].round(1).tolist())
