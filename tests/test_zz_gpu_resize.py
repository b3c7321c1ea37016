"""GPU: the patch resize of N1 (dm_resize_area) against the oracle restatement of cv2's INTER_AREA and the golden
outputs of the executed reference (ExtractFeatureDataset.resize_data, MyUtils2.py:362-376).  Bit exact.
"""
import os

import numpy as np
import pytest

from oracle.resize_area import resize_data

pytestmark = pytest.mark.gpu


def test_resize_windows_matches_the_executed_reference(cuda, golden_dir):
    import torch
    from deepmerge_b200.MyUtils2 import resize_windows
    g = np.load(os.path.join(golden_dir, "resize.npz"))
    for i, (s, t) in enumerate(g["cases"]):
        x = torch.from_numpy(g[f"in{i}"][None]).to(cuda)                 # [1, C, s, s]
        got = resize_windows(x, int(t))[0].cpu().numpy()
        assert got.dtype == np.float32 and np.array_equal(got, g[f"out{i}"]), (int(s), int(t))


def test_resize_windows_size_sweep_against_the_oracle(cuda):
    import torch
    from deepmerge_b200.MyUtils2 import resize_windows
    rng = np.random.default_rng(4)
    for t in (32, 64, 128, 1):
        for s in sorted(set(rng.integers(2, 3 * max(t, 40), 10).tolist() + [t, 2 * t, 3 * t, 4 * t])):
            x = rng.integers(0, 256, (3, 2, s, s), dtype=np.uint8)
            got = resize_windows(torch.from_numpy(x).to(cuda), t).cpu().numpy()
            want = np.stack([resize_data(w, t) for w in x])
            assert np.array_equal(got, want), (s, t)


def test_point_patches_equal_the_reference_loader_per_point(cuda):
    """get_patches_by_scales (MyUtils2.py:286-298) for all points at once: cut + resize + / 255 on the GPU."""
    import torch
    from deepmerge_b200 import MyUtils2
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (3, 180, 240), dtype=np.uint8)
    n = 12
    inner, obj = rng.integers(8, 40, n), rng.integers(40, 90, n)
    scales = np.stack([inner, obj, 2 * obj - inner, 3 * obj - 2 * inner], axis=1)
    w = {"ids": np.arange(n), "scales": scales, "xpix": rng.integers(1, 241, n), "ylin": rng.integers(1, 181, n)}
    got = MyUtils2.point_patches(torch.from_numpy(img).to(cuda), w)
    for k, cfg in enumerate(MyUtils2.SCALES):
        assert got[k].shape == (n, 3, cfg, cfg)
        for i in range(n):
            win = MyUtils2.calculate_left_top_point_and_size(int(w["xpix"][i]), int(w["ylin"][i]), int(scales[i, k]))
            want = resize_data(MyUtils2.cut_image(img, win), cfg)
            assert np.array_equal(got[k][i].cpu().numpy(), want), (i, k)
