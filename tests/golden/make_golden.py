"""Generate golden vectors by IMPORTING THE REFERENCE's own pure-torch code in the authoring container.

Run once where /root/reference exists:   python tests/golden/make_golden.py
It writes tests/golden/*.npz, which are committed; nothing at test time reads /root/reference.

Sources executed (unmodified, loaded by file path so that ``internal/__init__`` and its lightning imports are not needed):
  /root/reference/internal/utils/gaussian_projection.py   project_gaussians, build_gaussian_sort_key, build_tile_bounds
  /root/reference/internal/utils/sh_utils.py              eval_sh, eval_sh_decomposed
Fixture (1) is the literal known-answer input block of /root/reference/tests/gaussian_projection_test.py:30-63, stored
together with the expected values hard-coded at :97-113 of that file.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


gp = _load("ref_gaussian_projection", "internal/utils/gaussian_projection.py")
sh = _load("ref_sh_utils", "internal/utils/sh_utils.py")


def run_projection(means, scales, quats, w2c, fx, fy, cx, cy, H, W, cot_seed=1):
    means = means.clone().requires_grad_(True)
    scales = scales.clone().requires_grad_(True)
    quats = quats.clone().requires_grad_(True)
    out = gp.project_gaussians(
        means_3d=means, scales=scales, scale_modifier=1.0, quaternions=quats, world_to_camera=w2c,
        fx=torch.tensor(fx), fy=torch.tensor(fy), cx=torch.tensor(cx), cy=torch.tensor(cy),
        img_height=torch.tensor(H, dtype=torch.int), img_width=torch.tensor(W, dtype=torch.int), block_width=16)
    xys, depths, radii, conic, comp, tiles, cov3d, mask, rmin, rmax = out
    g = torch.Generator().manual_seed(cot_seed)
    c_xy = torch.randn(xys.shape, generator=g)
    c_d = torch.randn(depths.shape, generator=g)
    c_con = torch.randn(conic.shape, generator=g)
    c_comp = torch.randn(comp.shape, generator=g)
    loss = (xys * c_xy).sum() + (depths * c_d).sum() + (conic * c_con).sum() + (comp * c_comp).sum()
    loss.backward()
    return {
        "xys": xys, "depths": depths, "radii": radii, "conic": conic, "comp": comp, "tiles": tiles, "cov3d": cov3d,
        "mask": mask, "rect_min": rmin, "rect_max": rmax, "cot_xy": c_xy, "cot_depth": c_d, "cot_conic": c_con,
        "cot_comp": c_comp, "g_means": means.grad, "g_scales": scales.grad, "g_quats": quats.grad,
    }


def to_np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def kat():
    dtype = torch.float
    means = torch.tensor([[4.9744410514831543, -1.6869305372238159, -1.0178891420364380],
                          [0.1855451613664627, 0.2173379510641098, -1.6864157915115356],
                          [14.9114608764648438, -4.6346273422241211, 1.8997575044631958],
                          [5.0085635185241699, -3.8657102584838867, -1.3707503080368042]], dtype=dtype)
    scales = torch.tensor([[0.1152868643403053, 0.0463323593139648, 0.0125905377790332],
                           [0.0036764058750123, 0.0155582446604967, 0.0025763553567231],
                           [0.0729999020695686, 0.1261776685714722, 0.0579524375498295],
                           [1.8269745111465454, 0.1552953571081161, 0.2113087177276611]], dtype=dtype)
    quats = torch.tensor([[0.6251348853111267, -0.7321968674659729, 0.2666733860969543, 0.0444900505244732],
                          [0.9881987571716309, -0.0445680879056454, -0.1419259905815125, 0.0365220829844475],
                          [0.9662694931030273, 0.1446461081504822, -0.1685470491647720, 0.1303553283214569],
                          [0.8739961385726929, -0.3649578392505646, 0.1373531222343445, -0.2899493575096130]], dtype=dtype)
    W, H = 1297, 840
    fx, fy, cx, cy = 961.4099731445312500, 962.8024902343750000, 648.5, 420.0
    w2c = torch.tensor([
        [9.9991554021835327e-01, -1.2848137877881527e-02, -1.9360868027433753e-03, 0.0],
        [-5.9221056289970875e-04, -1.9391909241676331e-01, 9.8101717233657837e-01, 0.0],
        [-1.2979693710803986e-02, -9.8093330860137939e-01, -1.9391019642353058e-01, 0.0],
        [-3.2830274105072021e-01, -1.9259561300277710e+00, 3.9580578804016113e+00, 1.0]], dtype=dtype)
    out = run_projection(means, scales, quats, w2c, fx, fy, cx, cy, H, W)
    d = to_np(out)
    d.update(to_np({"means": means, "scales": scales, "quats": quats, "w2c": w2c}))
    d.update({"intr": np.array([fx, fy, cx, cy, W, H], dtype=np.float64)})
    # expected literals of the reference test (tests/gaussian_projection_test.py:97-113)
    d["expect_radii"] = np.array([0, 4, 0, 16783], dtype=np.int32)
    d["expect_tiles_masked"] = np.array([4, 4346], dtype=np.int32)
    d["expect_comp_masked"] = np.array([0.5893613696098328, 0.9999994039535522], dtype=np.float32)
    d["expect_conic_masked"] = np.array([[1.1229337453842163e+00, 1.4079402387142181e-01, 1.5783417224884033e+00],
                                         [2.3913329982860887e-07, -9.6377800673508318e-07, 4.5153879000281449e-06]], dtype=np.float32)
    d["expect_xys_masked_old_ndc_variant"] = np.array([[622.1340942382812500, 351.8106079101562500],
                                                      [11359.6181640625000000, 656.7397460937500000]], dtype=np.float32)
    d["expect_cov3d_upper_masked"] = np.array([
        [1.3772079910268076e-05, -1.3363457583182026e-05, 3.2048776574811200e-06, 2.3899228835944086e-04,
         -2.2861815523356199e-05, 9.4481683845515363e-06],
        [2.1180632114410400e+00, -1.5923748016357422e+00, -6.8420924246311188e-02, 1.2517973184585571e+00,
         6.5218225121498108e-02, 3.6743372678756714e-02]], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "kat_projection.npz"), **d)
    print("kat: radii", d["radii"], "tiles", d["tiles"], "comp", d["comp"])


def scene_case(name, n, W, H, seed, pose):
    from b200gs.scene import make_scene, activate, make_ring_cameras
    sc = activate(make_scene(n, seed, extent=1.3, mean_scale=0.05 if n <= 4096 else 0.01))
    cam = make_ring_cameras(W, H)[pose]
    out = run_projection(sc["means"], sc["scales"], sc["rotations"], cam.world_to_camera, float(cam.fx), float(cam.fy),
                         float(cam.cx), float(cam.cy), H, W)
    d = to_np(out)
    # SH through the reference's eval_sh ([N,3,K] layout) and its dc/rest variant, deg 0..3, with gradients
    dirs = sc["means"] - cam.camera_center
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    g = torch.Generator().manual_seed(7)
    cot = torch.randn(n, 3, generator=g)
    d["sh_cot"] = cot.numpy()
    for deg in range(4):
        shs = sc["shs"].clone().requires_grad_(True)
        dd = dirs.clone().requires_grad_(True)
        rgb = sh.eval_sh(deg, shs.transpose(1, 2), dd)
        rgb2 = sh.eval_sh_decomposed(deg, shs[:, :1, :], shs[:, 1:, :], dd)
        assert torch.allclose(rgb, rgb2, atol=1e-6)
        (torch.clamp_min(rgb + 0.5, 0.0) * cot).sum().backward()
        d[f"sh_rgb_deg{deg}"] = rgb.detach().numpy()
        d[f"sh_g_shs_deg{deg}"] = shs.grad.numpy()
        d[f"sh_g_dirs_deg{deg}"] = (dd.grad if dd.grad is not None else torch.zeros_like(dd)).numpy()
    # sort keys via the reference's python triple loop (small N only)
    if n <= 4096:
        tb = gp.build_tile_bounds(torch.tensor(H), torch.tensor(W), 16, "cpu")
        cums = torch.cumsum(out["tiles"], 0)
        keys, ids = gp.build_gaussian_sort_key(out["depths"].detach(), out["rect_min"], out["rect_max"], tb, cums)
        d["sort_key"] = keys.numpy()
        d["sort_ids"] = ids.numpy()
    d["meta"] = np.array([n, W, H, seed, pose], dtype=np.int64)
    if n > 4096:  # keep the committed fixture small: forward outputs only
        keep = {"xys", "depths", "radii", "conic", "comp", "tiles", "mask", "rect_min", "rect_max", "sh_rgb_deg3", "meta"}
        d = {k: v for k, v in d.items() if k in keep}
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)
    print(name, "V", int(out["mask"].sum()), "I", int(out["tiles"].sum()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    kat()
    scene_case("scene_n256_64x48", 256, 64, 48, 3, 0)
    scene_case("scene_n4096_256x256", 4096, 256, 256, 5, 3)
    scene_case("scene_n30000_800x800", 30000, 800, 800, 0, 0)
