"""BASELINE.json configs[4]: segment-count sweep 1k .. 4M on a fixed 16k x 16k raster (fragmentation stress
on the label cache, the atomics and the union-find).  Prints one JSON line per segment count with the
time of the fused raster pass, of the whole step, and size-independent checks (pixel count, perimeter /
boundary-length identity, label map idempotence)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmerge_b200 import MergeEngine, _lib
from deepmerge_b200.raster import _p, _stream
from deepmerge_b200.synth import synth_scene

side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
counts = [int(a) for a in sys.argv[2:]] or [1000, 4000, 16000, 64000, 256000, 1000000, 4000000]
L = _lib.lib()
dev = torch.device("cuda:0")
C = 4
for R_target in counts:
    sc = synth_scene(side, side, R_target, C=C, device=dev)
    R, N = sc.n_regions, sc.feats.shape[0]
    eng = MergeEngine(side, side, R, 100, C=C, n_points=N, device=dev)
    run = lambda: eng.run(sc.labels, sc.feats, 0.5, image=sc.image, xs=sc.xs, ys=sc.ys)
    res = run(); res = run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for _ in range(3):
        res = run()
    ev[1].record()
    for _ in range(3):
        eng.stats.zero_()
        L.check(L.dm_rag_scan(_p(sc.labels), side, side, side, side, _p(sc.image), C, side * C, R, 1, 1, _p(eng.area), _p(eng.border),
                              _p(eng.bsum), _p(eng.bsq), eng.cap, _p(eng.counts), _p(eng.ws), eng.ws_bytes, _stream()), "scan")
    ev[2].record()
    torch.cuda.synchronize()
    step_ms, rag_ms = ev[0].elapsed_time(ev[1]) / 3, ev[1].elapsed_time(ev[2]) / 3
    raw = int(eng.counts[1])
    g = eng.run(sc.labels, sc.feats, 0.5, image=sc.image, xs=sc.xs, ys=sc.ys, max_rounds=0, relabel=False)
    E0 = int(g.edge_keys.shape[0])
    area_ok = int(g.area.sum()) == side * side
    # every pixel side is either an image-border side or one half of a counted boundary pair
    perim_ok = int(g.perimeter.sum()) == 2 * int(g.boundary_len.to(torch.int64).sum()) + 4 * side
    res = run()
    roots = int((res.root == torch.arange(R, device=dev, dtype=torch.int32)).sum())
    again = torch.empty_like(res.labels)
    L.check(L.dm_relabel(_p(res.labels), side, side, side, _p(res.root), R, _p(again), side, _stream()), "relabel")
    idem = bool(torch.equal(again, res.labels))
    print(json.dumps({"side": side, "segments": R, "px_per_segment": side * side / R, "edges": E0, "raw_entries": raw,
                      "rag_ms": rag_ms, "rag_GBps": (4 + C) * side * side / rag_ms / 1e6, "step_ms": step_ms,
                      "Mpx_per_s": side * side / step_ms / 1e3, "segments_after": roots, "rounds": res.rounds,
                      "area_sum_ok": area_ok, "perimeter_identity_ok": perim_ok, "relabel_idempotent": idem}), flush=True)
    del eng, sc, res, g, again
    torch.cuda.empty_cache()
