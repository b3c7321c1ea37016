"""GPU: BASELINE.json configs[1] at FULL size (10k x 10k, 4 bands, ~100k segments) through size-independent
properties -- the oracle cannot run this size in seconds, so the checks are identities the domain offers:
pixel-count and checksum conservation, the perimeter / boundary-length identity, row-tile additivity, relabel
idempotence, run-to-run determinism, and agreement of the merged statistics with a recount of the final map."""
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
    # the merged statistics equal a recount of the final label map
    again = build_rag(labels_a, R, scene.image)
    assert torch.equal(again.area[roots], area_a[roots]) and int(again.area[~roots].sum()) == 0
    assert torch.equal(again.perimeter[roots], perim_a[roots])
    E = a.edge_keys.shape[0]
    assert torch.equal(again.edge_keys, a.edge_keys) and torch.equal(again.boundary_len, a.boundary_len[:E])
    # no surviving edge is below the threshold (the loop ran to its fixed point)
    assert a.rounds < 64 and bool((a.scores >= 0.5).all())
