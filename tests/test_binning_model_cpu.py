"""The ALGORITHM of the CUDA binning (csrc/binning.cu: stable depth sort -> stable partition of (cell, Gaussian) pairs with tile
masks -> order-preserving multi-split in 256-entry chunks) restated on CPU (oracle/bin_model.py) and pinned against the reference
order: the stable sort of the 64-bit (tile | depth bits) keys (oracle.gs_oracle, golden-pinned to the reference's own
build_gaussian_sort_key in test_oracle_golden.py).  The CUDA implementation itself is compared element for element with the
same oracle lists in tests/test_gpu_parity.py::test_binning_exact / test_binning_exact_stress."""
import pytest
import torch

from oracle import bin_model
from oracle import gs_oracle as O


def _scene(n, W, H, seed, pose, ms):
    from b200gs.scene import activate, make_ring_cameras, make_scene
    sc = activate(make_scene(n, seed, mean_scale=ms))
    cam = make_ring_cameras(W, H)[pose]
    ov = O.make_view(cam.R, cam.T, float(cam.fx), float(cam.fy), float(cam.cx), float(cam.cy), int(cam.width), int(cam.height))
    return sc, ov


# (n, W, H, seed, pose, mean_scale): one cell; several cells with partial edge cells; lists longer than one chunk; huge splats
CASES = [(256, 64, 48, 3, 0, 0.05), (600, 300, 200, 5, 2, 0.08), (1500, 272, 144, 7, 4, 0.02), (60, 400, 304, 9, 1, 1.0)]


@pytest.mark.parametrize("mode", [O.MODE_VANILLA, O.MODE_GSPLAT])
@pytest.mark.parametrize("n,W,H,seed,pose,ms", CASES)
def test_multisplit_equals_stable_sort(mode, n, W, H, seed, pose, ms):
    sc, ov = _scene(n, W, H, seed, pose, ms)
    ref = O.project(mode, sc["means"], sc["scales"], sc["rotations"], ov)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    keys, ids = O.build_sort_keys(ref["depth"], ref["rect_min"], ref["rect_max"], ref["tiles"], gx)
    _, sids, ranges = O.sort_and_ranges(keys, ids, gx * gy)
    got_ids, got_ranges = bin_model.hierarchical_binning(ref["depth"], ref["rect_min"], ref["rect_max"], gx, gy)
    assert got_ids.numel() == int(ref["tiles"].sum()) > 0
    assert torch.equal(got_ids, sids)
    nonempty = ranges[:, 1] > ranges[:, 0]
    assert torch.equal(got_ranges[nonempty], ranges[nonempty]) and int(got_ranges[~nonempty].abs().sum()) == 0


def test_multisplit_with_dropped_pairs_is_a_subsequence():
    """Tile culling only clears mask bits: every tile's list must be the order-preserving subsequence of the full list."""
    n, W, H = 800, 272, 208
    sc, ov = _scene(n, W, H, 13, 3, 0.06)
    ref = O.project(O.MODE_GSPLAT, sc["means"], sc["scales"], sc["rotations"], ov)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    keep = torch.rand(n, gy, gx, generator=torch.Generator().manual_seed(0)) < 0.6
    full_ids, full_ranges = bin_model.hierarchical_binning(ref["depth"], ref["rect_min"], ref["rect_max"], gx, gy)
    ids, ranges = bin_model.hierarchical_binning(ref["depth"], ref["rect_min"], ref["rect_max"], gx, gy, keep)
    assert 0 < ids.numel() < full_ids.numel()
    fi, ci = full_ids.tolist(), ids.tolist()
    for t in range(gx * gy):
        (fs, fe), (cs, ce) = full_ranges[t].tolist(), ranges[t].tolist()
        want = [g for g in fi[fs:fe] if bool(keep[g, t // gx, t % gx])]
        assert ci[cs:ce] == want
