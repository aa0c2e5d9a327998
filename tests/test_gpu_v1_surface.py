"""gsplat-v1 surface (b200gs/v1.py; reference: internal/renderers/gsplat_v1_renderer.py): hooks, multi-channel rasterization, exact
tile-based culling, absgrad / has_hit_any_pixels side channels — against the v0-surface renderer and the float64 oracle."""
import pytest
import torch

from oracle import gs_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(n=4096, W=256, H=192, seed=5, pose=3, ms=0.05):
    from b200gs.scene import SyntheticGaussians, make_ring_cameras, make_scene
    raw = make_scene(n, seed, mean_scale=ms)
    return raw, SyntheticGaussians(raw).to(DEV), make_ring_cameras(W, H)[pose].to_device(DEV)


def test_v1_rgb_matches_v0_surface_and_culling_is_exact():
    from b200gs.renderers import B200GSplatRenderer
    from b200gs.v1 import B200GSplatV1Renderer
    raw, model, cam = _setup()
    bg = torch.tensor([0.2, 0.1, 0.5], device=DEV)
    ref = B200GSplatRenderer(fused_activations=False).to(DEV)(cam, model, bg)
    outs = {}
    for cull in (False, True):
        r = B200GSplatV1Renderer(tile_based_culling=cull).instantiate().to(DEV)
        outs[cull] = r(cam, model, bg)
        # the v0-surface renderer normalises the quaternions once more (gsplat_renderer.py:68): equal up to that rounding
        assert float((outs[cull]["render"] - ref["render"]).abs().max()) < 2e-5
        assert int((outs[cull]["radii"] != ref["radii"]).sum()) <= 1 
        assert torch.equal(outs[cull]["visibility_filter"], outs[cull]["radii"] > 0)
        assert outs[cull]["viewspace_points"].shape == (raw["means"].shape[0], 2)
        assert torch.allclose(outs[cull]["viewspace_points_grad_scale"].cpu(), 0.5 * torch.tensor([[cam.width, cam.height]], dtype=torch.float32))
    assert torch.equal(outs[False]["render"], outs[True]["render"])               # exact culling: a subsequence of the lists, bit-identical image
    t_full, t_cull = outs[False]["isects"].binning.total, outs[True]["isects"].binning.total
    assert 0 < t_cull < t_full == int(outs[False]["isects"][0].sum())               # tiles_per_gauss = the rect pair count
    # has_hit_any_pixels: only visible splats, and every splat with a non-zero colour gradient did hit a pixel
    hit = outs[True]["acc_vis"]
    assert hit.dtype == torch.bool and bool((hit <= outs[True]["visibility_filter"]).all()) and int(hit.sum()) > 0
    assert torch.equal(hit, outs[False]["acc_vis"])


def test_v1_multichannel_side_channels_and_hooks():
    from b200gs.v1 import B200GSplatV1Renderer, B200GSplatV1RendererModule
    raw, model, cam = _setup()
    bg = torch.tensor([0.2, 0.1, 0.5], device=DEV)
    r = B200GSplatV1Renderer().instantiate().to(DEV)
    out = r(cam, model, bg, render_types=["rgb", "acc_depth", "normal", "alpha", "exp_depth", "inverse_depth", "hard_depth"])
    for k, c in (("render", 3), ("acc_depth", 1), ("normal", 3), ("alpha", 1), ("exp_depth", 1), ("inverse_depth", 1), ("hard_depth", 1)):
        assert out[k].shape == (c, int(cam.height), int(cam.width)) and bool(torch.isfinite(out[k]).all()), k
    # the 7 channels composited in two groups equal single renders of each feature
    rgb_only = r(cam, model, bg)
    assert torch.equal(out["render"], rgb_only["render"])
    depth_only = r(cam, model, bg, render_types=["acc_depth"])
    assert torch.allclose(out["acc_depth"], depth_only["acc_depth"], atol=1e-6)
    # alpha = 1 - T of the oracle
    act = {k: v.double() for k, v in __import__("b200gs.scene", fromlist=["activate"]).activate(raw).items()}
    from b200gs.scene import make_ring_cameras
    c = make_ring_cameras(int(cam.width), int(cam.height))[3]       # the same pose on the host (to_device moves a camera in place)
    ov = O.make_view(c.R, c.T, float(c.fx), float(c.fy), float(c.cx), float(c.cy), int(c.width), int(c.height))
    ref = O.render(O.MODE_GSPLAT, act["means"], act["scales"], act["rotations"], act["opacities"], act["shs"], ov, bg.cpu().double())
    assert float((out["alpha"][0].cpu().double() - ref["alpha"]).abs().max()) < 1e-4
    # absgrad: set by backward, >= |grad| elementwise (sum of absolute per-pixel contributions)
    out = r(cam, model, bg)
    vp = out["viewspace_points"]
    vp.retain_grad()
    out["render"].square().sum().backward()
    assert vp.grad is not None and hasattr(vp, "absgrad") and vp.absgrad.shape == vp.grad.shape
    assert bool((vp.absgrad >= vp.grad.abs() - 1e-4 * vp.absgrad.max()).all()) and float(vp.absgrad.sum()) > 0
    hit = out["acc_vis"]
    assert bool((vp.grad.abs().sum(dim=1)[~hit] == 0).all())
    # gradient parity of the v1 path with the oracle (rgb only)
    for p in model.parameters():
        p.grad = None
    ap = {k: v.clone().requires_grad_(True) for k, v in act.items()}
    cot = torch.rand(3, int(c.height), int(c.width), generator=torch.Generator().manual_seed(1)) * 2 - 1
    refg = O.render(O.MODE_GSPLAT, ap["means"], ap["scales"], ap["rotations"], ap["opacities"], ap["shs"], ov, bg.cpu().double())
    (refg["render"] * cot.double()).sum().backward()
    out = r(cam, model, bg)
    (out["render"] * cot.to(DEV)).sum().backward()
    g = model.gaussians["means"].grad.cpu().double()
    assert float((g - ap["means"].grad).abs().max() / ap["means"].grad.abs().max()) < 1e-3

    # hooks: a derived renderer that halves the opacities and paints everything white
    class Derived(B200GSplatV1RendererModule):
        def get_opacities(self, camera, gaussian_model, projections, visibility_filter, status, **kwargs):
            return gaussian_model.get_opacities().squeeze(-1) * 0.5, status

        def get_rgbs(self, camera, gaussian_model, projections, visibility_filter, status, **kwargs):
            return torch.ones(gaussian_model.get_means().shape[0], 3, device=DEV)

    d = Derived(B200GSplatV1Renderer()).to(DEV)
    od = d(cam, model, torch.zeros(3, device=DEV), render_types=["rgb", "alpha"])
    assert torch.allclose(od["render"][0], od["alpha"][0], atol=1e-6)            # white splats on black: every channel equals alpha
    assert float(od["alpha"].mean()) < float(out["alpha"].mean()) if out["alpha"] is not None else True


def test_v0_renderer_absgrad_option_matches_v1_side_channel():
    """`absgrad: true` of the density controller (vanilla_density_controller.py:112-113) reads `viewspace_points.absgrad`: the v0-surface
    renderer provides it through its `absgrad` option (also on the config dataclass); same values as the v1 renderer's side channel."""
    from b200gs.renderers import B200GSplatRenderer, B200GSplatRendererConfig
    from b200gs.v1 import B200GSplatV1Renderer
    raw, model, cam = _setup()
    bg = torch.tensor([0.2, 0.1, 0.5], device=DEV)
    r0 = B200GSplatRendererConfig(absgrad=True).instantiate().to(DEV)
    assert isinstance(r0, B200GSplatRenderer) and r0.absgrad
    out0 = r0(cam, model, bg)
    vp0 = out0["viewspace_points"]
    vp0.retain_grad()
    out0["render"].square().sum().backward()
    assert vp0.grad is not None and hasattr(vp0, "absgrad") and vp0.absgrad.shape[0] == vp0.grad.shape[0]
    assert bool((vp0.absgrad >= vp0.grad.abs()[:, :2] - 1e-4 * vp0.absgrad.max()).all()) and float(vp0.absgrad.sum()) > 0
    for p in model.parameters():
        p.grad = None
    r1 = B200GSplatV1Renderer(tile_based_culling=True).instantiate().to(DEV)
    out1 = r1(cam, model, bg)
    vp1 = out1["viewspace_points"]
    vp1.retain_grad()
    out1["render"].square().sum().backward()
    scale = float(vp1.absgrad.max())
    assert float((vp0.absgrad - vp1.absgrad).abs().max()) < 2e-4 * scale
    # without the option nothing is attached (and the fused single-node path stays in use)
    out2 = B200GSplatRenderer().to(DEV)(cam, model, bg)
    out2["render"].sum().backward()
    assert not hasattr(out2["viewspace_points"], "absgrad")
