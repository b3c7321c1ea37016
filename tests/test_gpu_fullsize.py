"""GPU: BASELINE.json configs[1] at FULL size (10k x 10k, 4 bands, ~100k segments).

test_whole_step_against_the_oracle_at_full_size checks the whole step -- RAG, band sums, pooled embeddings, scores,
merge loop, final label map -- bit for bit against the oracle (the plain-C restatement for the raster stages, which
takes a few seconds at this size, the numpy restatement for the graph stages).  The other tests are size-independent
identities the domain offers: pixel-count and checksum conservation, the perimeter / boundary-length identity,
row-tile additivity, relabel idempotence, run-to-run determinism."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H = W = 10000
R_TARGET, C = 100000, 4


@pytest.fixture(scope="module")
def scene(cuda):
    from deepmerge_b200.synth import synth_scene
    return synth_scene(H, W, R_TARGET, C=C, device=cuda)


def test_rag_conservation_identities(cuda, scene):
    import torch
    from deepmerge_b200 import build_rag
    rag = build_rag(scene.labels, scene.n_regions, scene.image)
    assert int(rag.area.sum()) == H * W                                          # every pixel counted once
    # checksum of checksums: the band sums of all regions add up to the band sums of the image
    img_sum = scene.image.view(-1, C).to(torch.int64).sum(0)
    assert torch.equal(rag.band_sum.sum(0), img_sum)
    img_sq = (scene.image.view(-1, C).to(torch.int64) ** 2).sum(0)
    assert torch.equal(rag.band_sumsq.sum(0), img_sq)
    # every pixel side is an image-border side or one half of a counted neighbour pair
    blen = rag.boundary_len.to(torch.int64)
    assert int(rag.perimeter.sum()) == 2 * int(blen.sum()) + 2 * (H + W)
    assert int(rag.border.sum()) == 2 * (H + W)
    k = rag.edge_keys
    assert bool((k[1:] > k[:-1]).all())                                          # sorted, unique
    lo, hi = rag.endpoints()
    assert bool((lo < hi).all()) and int(hi.max()) < scene.n_regions
    # the number of differing neighbour pairs, recounted straight from the raster
    L = scene.labels
    pairs = int((L[:, 1:] != L[:, :-1]).sum()) + int((L[1:] != L[:-1]).sum())
    assert int(blen.sum()) == pairs


def test_row_tiles_add_up_to_the_whole_scene(cuda, scene):
    import torch
    from deepmerge_b200 import build_rag, merge_edge_lists
    whole = build_rag(scene.labels, scene.n_regions, scene.image)
    cuts = [0, 3333, 7000, H]
    ks, bs, area, border, bsum = [], [], 0, 0, 0
    for y0, y1 in zip(cuts[:-1], cuts[1:]):
        last = y1 == H
        r = build_rag(scene.labels[y0:y1 + (0 if last else 1)], scene.n_regions, scene.image[y0:y1 + (0 if last else 1)],
                      rows_own=y1 - y0, top_border=y0 == 0, bottom_border=last)
        ks.append(r.edge_keys); bs.append(r.boundary_len)
        area, border, bsum = area + r.area, border + r.border, bsum + r.band_sum
    k, b = merge_edge_lists(torch.cat(ks), torch.cat(bs), scene.n_regions)
    assert torch.equal(k, whole.edge_keys) and torch.equal(b, whole.boundary_len)
    assert torch.equal(area, whole.area) and torch.equal(border, whole.border) and torch.equal(bsum, whole.band_sum)


def test_merge_is_deterministic_idempotent_and_consistent(cuda, scene):
    import torch
    from deepmerge_b200 import MergeEngine, build_rag, relabel
    eng = MergeEngine(H, W, scene.n_regions, scene.feats.shape[1], C=C, n_points=scene.feats.shape[0], device=cuda)
    run = lambda: eng.run(scene.labels, scene.feats, 0.5, image=scene.image, xs=scene.xs, ys=scene.ys)
    a = run()
    labels_a, root_a, area_a, perim_a = a.labels.clone(), a.root.clone(), a.area.clone(), a.perimeter.clone()
    b = run()
    assert torch.equal(labels_a, b.labels) and torch.equal(root_a, b.root) and a.merges == b.merges    # bit-equal reruns
    R = scene.n_regions
    ids = torch.arange(R, device=cuda, dtype=torch.int32)
    roots = root_a == ids
    assert torch.equal(root_a[root_a.long()], root_a)                            # roots are fixed points
    assert bool((root_a <= ids).all())                                           # root = minimum id of the component
    assert int(roots.sum()) == R - a.merges                                      # every merge removes one region
    assert torch.equal(relabel(labels_a, root_a), labels_a)                      # idempotent
    # the merged statistics equal those of the final label map rebuilt from scratch
    again = build_rag(labels_a, R, scene.image)
    assert torch.equal(again.area[roots], area_a[roots]) and int(again.area[~roots].sum()) == 0
    assert torch.equal(again.perimeter[roots], perim_a[roots])
    E = a.edge_keys.shape[0]
    assert torch.equal(again.edge_keys, a.edge_keys) and torch.equal(again.boundary_len, a.boundary_len[:E])
    # no surviving edge is below the threshold (the loop ran to its fixed point)
    # (a region without sample points has a NaN embedding: its edges score NaN and are never selected)
    assert a.rounds < 64 and not bool((a.scores < 0.5).any())


def test_whole_step_against_the_oracle_at_full_size(cuda, scene):
    """The BASELINE configuration itself, not a scaled-down stand-in: every integer output of the step equals the
    oracle's on the same inputs (the device-generated scene is copied to the host and fed to both)."""
    import torch
    from deepmerge_b200 import MergeEngine, build_rag
    from oracle import build as oc
    from oracle import oracle_np as o
    R = scene.n_regions
    L = scene.labels.cpu().numpy()
    img = scene.image.cpu().numpy()
    # --- raster stages: plain C restatement
    k, b, area, per = oc.build_rag(L, R)
    s, q = oc.pool_bands(L, img, R)
    rag = build_rag(scene.labels, R, scene.image)
    assert np.array_equal(rag.edge_keys.cpu().numpy().view(np.uint64), k)
    assert np.array_equal(rag.boundary_len.cpu().numpy().view(np.uint32), b)
    assert np.array_equal(rag.area.cpu().numpy(), area) and np.array_equal(rag.perimeter.cpu().numpy(), per)
    assert np.array_equal(rag.band_sum.cpu().numpy().view(np.uint64), s)
    assert np.array_equal(rag.band_sumsq.cpu().numpy().view(np.uint64), q)
    del rag, s, q, img
    # --- graph stages: numpy restatement (membership, np.mean-order pooling, L2 scores, merge loop)
    rop = L[scene.ys.cpu().numpy(), scene.xs.cpu().numpy()]
    assert np.array_equal(scene.region_of_point.cpu().numpy(), rop)
    feats = scene.feats.cpu().numpy()
    off, ids = o.csr_from_region_of_point(rop, R)
    ps, cnt, _ = o.pool_points_csr(off, ids, feats)
    want = o.merge_graph(ps, cnt, area, per, k, b, tau=0.5)
    want_labels = oc.relabel(L, want["root"])
    eng = MergeEngine(H, W, R, feats.shape[1], C=C, n_points=feats.shape[0], device=cuda)
    got = eng.run(scene.labels, scene.feats, 0.5, image=scene.image, xs=scene.xs, ys=scene.ys)
    assert got.rounds == want["rounds"] and got.merges == want["merges"]
    assert np.array_equal(got.root.cpu().numpy(), want["root"])
    assert np.array_equal(got.labels.cpu().numpy(), want_labels)                  # final label map, 1e8 pixels, bit exact
    roots = want["root"] == np.arange(R)
    assert np.array_equal(got.area.cpu().numpy()[roots], want["area"][roots])
    assert np.array_equal(got.perimeter.cpu().numpy()[roots], want["perim"][roots])
    E = len(want["keys"])
    assert np.array_equal(got.edge_keys.cpu().numpy().view(np.uint64), want["keys"])
    assert np.array_equal(got.boundary_len.cpu().numpy().view(np.uint32)[:E], want["blen"])
    assert np.array_equal(got.cnt.cpu().numpy()[roots], want["cnt"][roots])
    np.testing.assert_allclose(got.scores.cpu().numpy()[:E], want["scores"], rtol=1e-3, atol=1e-4)
