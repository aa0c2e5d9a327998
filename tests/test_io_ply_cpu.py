"""3DGS .ply reader/writer (b200gs/io_ply.py; reference: internal/utils/gaussian_utils.py:51-255): round trip, the channel-major
f_rest layout, name-based lookup (files with normals / colours / another property order), SH-degree detection."""
import numpy as np
import pytest
import torch


def _params(n=37, deg=3, seed=0):
    from b200gs.scene import make_scene
    return make_scene(n, seed, sh_degree=deg)


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_round_trip_and_layout(tmp_path, deg):
    from b200gs.io_ply import load_ply, save_ply
    p = _params(deg=deg)
    path = str(tmp_path / "pc.ply")
    save_ply(path, p, with_normals=(deg % 2 == 0), colors=torch.randint(0, 255, (37, 3), dtype=torch.uint8) if deg == 3 else None)
    q = load_ply(path)
    assert q["sh_degree"] == deg
    for k in ("means", "shs_dc", "shs_rest", "opacities", "scales", "rotations"):
        assert q[k].dtype == torch.float32 and q[k].shape == p[k].shape and torch.equal(q[k], p[k]), k
    # the on-disk f_rest order is channel-major: f_rest_j = features_rest[:, c, k] with j = c * (K-1) + k  (gaussian_utils.py:66,198-200)
    if deg > 0:
        from b200gs.io_ply import _read_vertex_table
        t = _read_vertex_table(path)
        km1 = (deg + 1) ** 2 - 1
        for c in range(3):
            for k in (0, km1 - 1):
                assert np.array_equal(t[f"f_rest_{c * km1 + k}"], p["shs_rest"][:, k, c].numpy())
        with pytest.raises(ValueError):
            load_ply(path, sh_degree=deg - 1)


def test_reads_any_property_order(tmp_path):
    """A file whose properties come in another order (and with extra ones) loads by name."""
    from b200gs.io_ply import load_ply
    n = 5
    rng = np.random.default_rng(3)
    names = ["rot_3", "rot_2", "rot_1", "rot_0", "scale_2", "scale_1", "scale_0", "opacity", "f_dc_2", "f_dc_1", "f_dc_0", "z", "y", "x", "confidence"]
    table = np.empty(n, dtype=[(k, "<f4") for k in names])
    for k in names:
        table[k] = rng.standard_normal(n).astype(np.float32)
    path = str(tmp_path / "odd.ply")
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment made by hand\nelement vertex %d\n" % n).encode())
        f.write("".join(f"property float {k}\n" for k in names).encode())
        f.write(b"end_header\n")
        table.tofile(f)
    q = load_ply(path)
    assert q["sh_degree"] == 0 and q["shs_rest"].shape == (n, 0, 3)
    assert np.array_equal(q["means"][:, 1].numpy(), table["y"]) and np.array_equal(q["rotations"][:, 3].numpy(), table["rot_3"])
    assert np.array_equal(q["shs_dc"][:, 0, 2].numpy(), table["f_dc_2"]) and np.array_equal(q["opacities"][:, 0].numpy(), table["opacity"])
