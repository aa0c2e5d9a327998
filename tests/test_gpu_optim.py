"""Visibility-masked Adam and densification statistics (csrc/optim.cu) against plain torch restatements of the reference-side formulas
(gsplat SelectiveAdam's update rule; vanilla_density_controller.py:101-123)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_selective_adam_updates_only_visible_rows():
    from b200gs.optimizers import B200SelectiveAdam
    g = torch.Generator().manual_seed(3)
    n = 5000
    shapes = {"means": (n, 3), "shs_rest": (n, 15, 3), "opacities": (n, 1), "rotations": (n, 4)}
    params = {k: torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for k, s in shapes.items()}
    lrs = {"means": 1.6e-4, "shs_rest": 1.25e-4, "opacities": 5e-2, "rotations": 1e-3}
    opt = B200SelectiveAdam().instantiate([{"params": [p], "name": k, "lr": lrs[k]} for k, p in params.items()], lr=1e-3)
    ref_p = {k: p.detach().cpu().clone().double() for k, p in params.items()}
    ref_m = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    ref_v = {k: torch.zeros_like(v) for k, v in ref_p.items()}
    b1, b2, eps = 0.9, 0.999, 1e-15
    for step in range(4):
        vis = torch.rand(n, generator=g) < 0.4
        grads = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
        for k, p in params.items():
            p.grad = grads[k].to(DEV)
        vp = torch.zeros(n, 2, device=DEV)
        vp.has_hit_any_pixels = vis.to(DEV)
        opt.on_after_backward({"viewspace_points": vp, "visibility_filter": torch.ones(n, dtype=torch.bool, device=DEV)}, None, None, step, None)
        opt.step()
        for k in shapes:
            gk = grads[k].double()
            # the kernel's (and gsplat's) constants are fp32: 1.0f - 0.999f = 0.00100004673 (4.7e-5 off the real 0.001)
            f32 = torch.float32
            b1f, b2f = float(torch.tensor(b1, dtype=f32)), float(torch.tensor(b2, dtype=f32))
            omb1, omb2 = float(torch.tensor(1.0, dtype=f32) - torch.tensor(b1, dtype=f32)), float(torch.tensor(1.0, dtype=f32) - torch.tensor(b2, dtype=f32))
            m = b1f * ref_m[k] + omb1 * gk
            v = b2f * ref_v[k] + omb2 * gk * gk
            upd = ref_p[k] - lrs[k] * m / (v.sqrt() + eps)
            sel = vis.cpu().reshape((n,) + (1,) * (gk.dim() - 1))
            ref_m[k], ref_v[k], ref_p[k] = torch.where(sel, m, ref_m[k]), torch.where(sel, v, ref_v[k]), torch.where(sel, upd, ref_p[k])
    for k, p in params.items():
        st = opt.state[p]
        assert torch.allclose(p.detach().cpu().double(), ref_p[k], rtol=2e-5, atol=1e-6), k
        assert torch.allclose(st["exp_avg"].cpu().double(), ref_m[k], rtol=1e-5, atol=1e-7), k
        assert torch.allclose(st["exp_avg_sq"].cpu().double(), ref_v[k], rtol=1e-5, atol=1e-9), k


def test_densification_stats_match_the_controller_formulas():
    from b200gs.optimizers import update_densification_stats
    g = torch.Generator().manual_seed(5)
    n = 7000
    for stride, scale in ((2, torch.tensor([[400.0, 300.0]])), (3, None)):
        radii = (torch.randint(0, 40, (n,), generator=g) * (torch.rand(n, generator=g) < 0.6)).to(torch.int32)
        grad = torch.randn(n, stride, generator=g)
        max_r, acc, den = torch.rand(n, generator=g) * 30, torch.rand(n, 1, generator=g), torch.randint(0, 5, (n, 1), generator=g).float()
        vis = radii > 0
        # reference formulas (vanilla_density_controller.py:107-123)
        r_max = max_r.clone()
        r_max[vis] = torch.max(r_max[vis], radii[vis].float())
        sg = grad[vis, :2] * (scale if scale is not None else 1.0)
        r_acc, r_den = acc.clone(), den.clone()
        r_acc[vis] += torch.norm(sg, dim=-1, keepdim=True)
        r_den[vis] += 1
        d = [t.to(DEV) for t in (max_r, acc, den)]
        update_densification_stats(radii.to(DEV), grad.to(DEV), d[0], d[1], d[2], visibility_filter=vis.to(DEV) if stride == 2 else None,
                                   scale=scale.to(DEV) if scale is not None else None)
        assert torch.equal(d[0].cpu(), r_max) and torch.allclose(d[1].cpu(), r_acc, rtol=1e-6, atol=1e-6) and torch.equal(d[2].cpu(), r_den)


@pytest.mark.parametrize("n,kind", [(3, "uniform"), (2000, "uniform"), (8000, "clustered"), (5000, "planar"), (4000, "duplicates")])
def test_knn_mean_dist2_matches_brute_force(n, kind):
    """simple_knn's distCUDA2: mean squared distance to the 3 nearest neighbours, exact, on distributions that stress the hash grid."""
    from b200gs import ops
    g = torch.Generator().manual_seed(n)
    if kind == "uniform":
        pts = torch.rand(n, 3, generator=g) * 4 - 2
    elif kind == "clustered":      # SfM-like: dense clumps + sparse outliers, very non-uniform cell occupancy
        centers = torch.randn(20, 3, generator=g) * 5
        pts = centers[torch.randint(0, 20, (n,), generator=g)] + 0.05 * torch.randn(n, 3, generator=g)
        pts[: n // 50] = torch.randn(n // 50, 3, generator=g) * 40
    elif kind == "planar":
        pts = torch.rand(n, 3, generator=g)
        pts[:, 2] = 0.25
    else:
        pts = torch.rand(n // 2, 3, generator=g).repeat(2, 1)
    out = ops.knn_mean_dist2(pts.to(DEV)).cpu()
    d = torch.cdist(pts.double(), pts.double()) ** 2
    d.fill_diagonal_(float("inf"))
    ref = d.topk(min(3, n - 1), largest=False).values.sum(dim=1) / 3.0
    assert torch.allclose(out.double(), ref, rtol=1e-4, atol=1e-9), float((out.double() - ref).abs().max())
    import b200gs.compat as compat
    compat.install()
    from simple_knn._C import distCUDA2
    assert torch.equal(distCUDA2(pts.to(DEV)).cpu(), out)
