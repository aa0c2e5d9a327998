"""gsplat v0-compat operator surface backed by libb200gs.so: the three functions the reference's GSPlatRenderer
calls (``internal/renderers/gsplat_renderer.py:2-4,64-106``)."""
from ..ops import project_gaussians, rasterize_gaussians, spherical_harmonics  # noqa: F401
