"""Deterministic synthetic scenes and camera rings — the benchmark/test input contract.

Scene law G(N, seed, extent, s): one ``torch.Generator`` seeded with ``seed``, draws in this order (fixes the visible
count V and the tile-pair count I exactly): ``rand(N,3)`` means in [-extent, extent]^3 -> ``randn(N,3)`` log-scale noise
(log s + 0.5 n) -> ``randn(N,4)`` quaternions -> ``randn(N,1)`` opacity logits (x1.5) -> ``randn(N,1,3)`` SH dc (x0.5)
-> ``randn(N,15,3)`` SH rest (x0.1).  Parameters are RAW (pre-activation), shaped like
``VanillaGaussianModel``'s ParameterDict (``internal/models/vanilla_gaussian.py:66``).

Cameras: pinhole, fov_x 39.6 deg, fy = fx, principal point at the image centre, 32 poses on a circle of radius 4
around the origin (pose 0: identity rotation, T = (0, 0, 4); pose k rotated about world Y by 2*pi*k/32).
"""
import math
from typing import Dict, List

import torch

from .cameras import Camera, make_camera

SH_COEFFS_DEG3 = 16


def make_scene(n: int, seed: int = 0, extent: float = 1.3, mean_scale: float = 0.01, sh_degree: int = 3) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    k_rest = (sh_degree + 1) ** 2 - 1
    means = (torch.rand(n, 3, generator=g) * 2.0 - 1.0) * extent
    scales = math.log(mean_scale) + 0.5 * torch.randn(n, 3, generator=g)
    rotations = torch.randn(n, 4, generator=g)
    opacities = 1.5 * torch.randn(n, 1, generator=g)
    shs_dc = 0.5 * torch.randn(n, 1, 3, generator=g)
    shs_rest = 0.1 * torch.randn(n, 15, 3, generator=g)
    if k_rest > 15:   # degree 4: drawn after everything else, so the degree <= 3 scenes keep their law
        shs_rest = torch.cat([shs_rest, 0.1 * torch.randn(n, k_rest - 15, 3, generator=g)], dim=1)
    shs_rest = shs_rest[:, :k_rest, :].contiguous()
    return {"means": means, "scales": scales, "rotations": rotations, "opacities": opacities, "shs_dc": shs_dc,
            "shs_rest": shs_rest}


def activate(scene: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The activations of VanillaGaussianModel's getters (vanilla_gaussian.py:345-358, gaussian.py:250-254)."""
    return {
        "means": scene["means"],
        "scales": torch.exp(scene["scales"]),
        "rotations": torch.nn.functional.normalize(scene["rotations"]),
        "opacities": torch.sigmoid(scene["opacities"]),
        "shs": torch.cat((scene["shs_dc"], scene["shs_rest"]), dim=1).contiguous(),
    }


def ring_pose(k: int, n_poses: int = 32, distance: float = 4.0):
    th = 2.0 * math.pi * k / n_poses
    c, s = math.cos(th), math.sin(th)
    R = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float32)
    T = torch.tensor([0.0, 0.0, distance], dtype=torch.float32)
    return R, T


def make_ring_cameras(width: int, height: int, n_poses: int = 32, distance: float = 4.0, fov_x_deg: float = 39.6) -> List[Camera]:
    fx = 0.5 * width / math.tan(math.radians(fov_x_deg) * 0.5)
    cams = []
    for k in range(n_poses):
        R, T = ring_pose(k, n_poses, distance)
        cams.append(make_camera(R, T, fx, fx, width / 2.0, height / 2.0, width, height, idx=k))
    return cams


class SyntheticGaussians(torch.nn.Module):
    """Parameter container with the getters the reference's renderers read
    (``VanillaGaussianModel``: internal/models/vanilla_gaussian.py:345-358,422-440; internal/models/gaussian.py:250-254)."""

    def __init__(self, scene: Dict[str, torch.Tensor], sh_degree: int = 3, active_sh_degree: int = None):
        super().__init__()
        self.gaussians = torch.nn.ParameterDict({k: torch.nn.Parameter(v.clone()) for k, v in scene.items()})
        self.max_sh_degree = sh_degree
        self.active_sh_degree = sh_degree if active_sh_degree is None else active_sh_degree

    @property
    def get_xyz(self):
        return self.gaussians["means"]

    @property
    def get_scaling(self):
        return torch.exp(self.gaussians["scales"])

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self.gaussians["rotations"])

    @property
    def get_opacity(self):
        return torch.sigmoid(self.gaussians["opacities"])

    @property
    def get_features(self):
        return torch.cat((self.gaussians["shs_dc"], self.gaussians["shs_rest"]), dim=1)

    def get_shs(self):
        return self.get_features

    # the method-style getters of the reference's current GaussianModel API (internal/models/gaussian.py:122-254), read by
    # the gsplat-v1 renderers
    is_pre_activated = False

    def get_means(self):
        return self.gaussians["means"]

    def get_scales(self):
        return self.get_scaling

    def get_rotations(self):
        return self.get_rotation

    def get_opacities(self):
        return self.get_opacity

    def get_shs_dc(self):
        return self.gaussians["shs_dc"]

    def get_shs_rest(self):
        return self.gaussians["shs_rest"]
