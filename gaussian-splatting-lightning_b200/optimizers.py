"""Visibility-masked Adam and the density controller's per-step statistics on the b200gs kernels.

``B200SelectiveAdam`` is selectable from YAML where the reference offers ``SelectiveAdam`` / ``SparseGaussianAdam``
(internal/optimizers.py:26-90): only the Gaussians that took part in the rendered view are stepped — their parameters and both
Adam moments; everything else is left untouched (no decay of the moments of invisible Gaussians, exactly like those CUDA
optimizers).  The arithmetic is theirs (no bias correction):  m = b1 m + (1-b1) g,  v = b2 v + (1-b2) g^2,
p -= lr m / (sqrt(v) + eps).

``update_densification_stats`` is ``VanillaDensityControllerImpl.update_states`` (vanilla_density_controller.py:101-123) in one
kernel.
"""
from dataclasses import dataclass
from typing import Tuple

import torch

from . import ops
from ._lib import check, lib, ptr


class SelectiveAdam(torch.optim.Optimizer):
    """``gsplat.optimizers.SelectiveAdam(params, eps, betas)``: ``step(visibility)`` with ``visibility`` bool [N]; every parameter
    tensor is [N, ...] (one row per Gaussian)."""

    def __init__(self, params, eps: float = 1e-15, betas: Tuple[float, float] = (0.9, 0.999), lr: float = 1e-3):
        super().__init__(params, dict(lr=lr, eps=eps, betas=betas))

    @torch.no_grad()
    def step(self, visibility: torch.Tensor):
        L = lib()
        vis = visibility.to(torch.uint8).contiguous()
        n = vis.numel()
        stream = ops._stream()
        for group in self.param_groups:
            lr, eps = float(group["lr"]), float(group["eps"])
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.shape[0] != n:
                    raise ValueError(f"parameter with {p.shape[0]} rows, visibility with {n}")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ValueError("SelectiveAdam: parameters must be contiguous float32 CUDA tensors")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                grad = p.grad.contiguous()
                check(L.b200gs_selective_adam(n, p.numel() // max(n, 1), ptr(p), ptr(grad), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), ptr(vis),
                                              lr, float(b1), float(b2), eps, stream), "b200gs_selective_adam")


@dataclass
class B200SelectiveAdam:
    """Optimizer config in the shape of the reference's (internal/optimizers.py:26-57): ``instantiate(params, lr)`` returns an optimizer
    whose ``on_after_backward`` hook picks up ``outputs["viewspace_points"].has_hit_any_pixels`` (or ``visibility_filter``) and
    whose ``step()`` uses it."""
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-15

    def instantiate(self, params, lr: float, *args, **kwargs):
        for group in params:
            if "lr" not in group:
                group["lr"] = lr

        class Adapter(SelectiveAdam):
            def on_after_backward(self, outputs, batch, gaussian_model, global_step, pl_module):
                hits = getattr(outputs["viewspace_points"], "has_hit_any_pixels", None)
                self.visibility = hits if hits is not None else outputs["visibility_filter"]

            def step(self, closure=None):
                loss = None
                if closure is not None:
                    with torch.enable_grad():
                        loss = closure()
                super().step(self.visibility)
                return loss

        return Adapter(params, eps=self.eps, betas=self.betas, lr=lr)


@torch.no_grad()
def update_densification_stats(radii: torch.Tensor, grad: torch.Tensor, max_radii2D: torch.Tensor, xyz_gradient_accum: torch.Tensor,
                               denom: torch.Tensor, visibility_filter: torch.Tensor = None, scale=None) -> None:
    """In place, on the visible Gaussians: max_radii2D = max(max_radii2D, radii); xyz_gradient_accum += |grad[:, :2] * scale|; denom += 1.
    radii int32 [N]; grad [N, >=2] (``viewspace_points.grad`` or ``.absgrad``); scale: None, a number or ``viewspace_points_grad_scale``
    ([1,2] / [2]); visibility_filter bool [N] (None: radii > 0)."""
    if scale is None:
        sx = sy = 1.0
    elif isinstance(scale, (int, float)):
        sx = sy = float(scale)
    else:
        s = scale.reshape(-1).tolist()
        sx, sy = (s[0], s[1]) if len(s) >= 2 else (s[0], s[0])
    n = radii.shape[0]
    vis = None if visibility_filter is None else visibility_filter.to(torch.uint8).contiguous()
    g = grad if grad.is_contiguous() else grad.contiguous()
    check(lib().b200gs_densify_stats(n, ptr(radii.contiguous()), ptr(vis), ptr(g), int(g.shape[1]), float(sx), float(sy), ptr(max_radii2D),
                                     ptr(xyz_gradient_accum), ptr(denom), ops._stream()), "b200gs_densify_stats")
