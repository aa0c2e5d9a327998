"""``diff_gaussian_rasterization`` look-alike backed by libb200gs.so.

Same names, argument meaning and error behaviour as the operator surface used at
``internal/renderers/vanilla_renderer.py:62-77,111-120``: ``GaussianRasterizationSettings`` (NamedTuple) and
``GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
rotations=None, cov3D_precomp=None) -> (color [3,H,W], radii int32 [N])``.
"""
from typing import NamedTuple

import torch

from .. import ops
from .._lib import MODE_VANILLA


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def settings_to_view(rs: GaussianRasterizationSettings, sh_stride: int = 1):
    return ops.make_view(MODE_VANILLA, rs.image_width, rs.image_height, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
                         viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, campos=rs.campos, sh_degree=int(rs.sh_degree),
                         sh_stride=sh_stride, scale_modifier=rs.scale_modifier)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings, view=None):
        super().__init__()
        self.raster_settings = raster_settings
        self._view = view

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if cov3D_precomp is not None:
            raise NotImplementedError("b200gs: cov3D_precomp is not supported (compute_cov3D_python defaults to False, "
                                      "vanilla_renderer.py:19)")
        view = self._view if self._view is not None else settings_to_view(rs)
        return ops.rasterize_vanilla(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, rs.bg, view)
