"""Renderer plug-ins for gaussian-splatting-lightning backed by libb200gs.so.

Select in YAML exactly like any other renderer of the reference (``configs/gsplat.yaml:1-2``):

    model:
      renderer: b200gs.renderers.B200VanillaRenderer      # or b200gs.renderers.B200GSplatRenderer

``B200VanillaRenderer`` mirrors ``internal/renderers/vanilla_renderer.py:18-213`` (same constructor arguments,
``forward`` signature, static ``render`` and return dict: render / viewspace_points / visibility_filter / radii).
``B200GSplatRenderer`` mirrors ``internal/renderers/gsplat_renderer.py:11-391`` (rgb, alpha and depth variants, statics
``render`` / ``project`` / ``rasterize`` / ``rasterize_simplified``; return dict incl. viewspace_points_grad_scale).

When the reference package is importable the plug-ins subclass ITS ``Renderer`` so ``isinstance`` checks in the
training loop hold; otherwise a local mirror of ``internal/renderers/renderer.py:10-117`` is used (tests, bench).
"""
import math
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional, Tuple

import torch

from . import ops
from ._lib import MODE_GSPLAT, MODE_VANILLA

try:  # inside the reference repo
    from internal.renderers.renderer import Renderer, RendererConfig, RendererOutputInfo, RendererOutputTypes  # type: ignore
except Exception:  # standalone mirror of internal/renderers/renderer.py
    class RendererOutputTypes:
        RGB: int = 1
        GRAY: int = 2
        NORMAL_MAP: int = 3
        FEATURE_MAP: int = 4
        OTHER: int = 65535

    @dataclass
    class RendererOutputInfo:
        key: str
        type: int = RendererOutputTypes.RGB
        visualizer: Callable = None

        def __post_init__(self):
            if self.type == RendererOutputTypes.OTHER and self.visualizer is None:
                raise ValueError("Visualizer must be provided when `type` is `OTHER`")

    class Renderer(torch.nn.Module):
        def forward(self, viewpoint_camera, pc, bg_color: torch.Tensor, scaling_modifier=1.0, render_types: list = None, **kwargs):
            pass

        def training_forward(self, step: int, module, viewpoint_camera, pc, bg_color: torch.Tensor, render_types: list = None, **kwargs):
            return self(viewpoint_camera=viewpoint_camera, pc=pc, bg_color=bg_color, render_types=render_types, **kwargs)

        def before_training_step(self, step: int, module):
            return

        def after_training_step(self, step: int, module):
            return

        def setup(self, stage: str, *args: Any, **kwargs: Any) -> Any:
            pass

        def training_setup(self, module) -> Tuple[Optional[Any], Optional[Any]]:
            return None, None

        def on_load_checkpoint(self, module, checkpoint):
            pass

        def setup_web_viewer_tabs(self, viewer, server, tabs):
            pass

        def get_available_outputs(self) -> Dict[str, RendererOutputInfo]:
            return {"rgb": RendererOutputInfo("render")}

    @dataclass
    class RendererConfig:
        def instantiate(self, *args, **kwargs) -> Renderer:
            raise NotImplementedError()


DEFAULT_BLOCK_SIZE: int = 16
DEFAULT_ANTI_ALIASED_STATUS: bool = True


# ----------------------------------------------------------------------------------------------------------------------
# camera -> host view struct, cached on the camera object (one D2H of the matrices per camera, not per step; the
# reference pays .item()/float() syncs every step: vanilla_renderer.py:59-60, gsplat_renderer.py:61-74)
# ----------------------------------------------------------------------------------------------------------------------
def camera_view(camera, mode: int, cache: bool = True):
    store = getattr(camera, "_b200gs_views", None) if cache else None
    if store is not None and mode in store:
        return store[mode]
    W, H = int(camera.width), int(camera.height)
    if mode == MODE_VANILLA:
        view = ops.make_view(MODE_VANILLA, W, H, tanfovx=math.tan(float(camera.fov_x) * 0.5),
                             tanfovy=math.tan(float(camera.fov_y) * 0.5), viewmatrix=camera.world_to_camera,
                             projmatrix=camera.full_projection, campos=camera.camera_center)
    else:
        view = ops.make_view(MODE_GSPLAT, W, H, fx=float(camera.fx), fy=float(camera.fy), cx=float(camera.cx),
                             cy=float(camera.cy), viewmatrix=camera.world_to_camera, campos=camera.camera_center)
    if cache:
        try:
            if store is None:
                store = {}
                setattr(camera, "_b200gs_views", store)
            store[mode] = view
        except Exception:
            pass
    return view


def _view_with(view, **kw):
    return ops._copy_view(view, **kw)


_FUSABLE_MODELS = ("VanillaGaussianModel", "SyntheticGaussians")
_RAW_KEYS = ("means", "scales", "rotations", "opacities", "shs_dc", "shs_rest")


def _raw_parameters(pc):
    """The raw parameter tensors of a vanilla Gaussian model, or None when the model is anything else (a subclass may
    override an activation, e.g. mip-splatting's filtered scales/opacities: those must go through the getters)."""
    if type(pc).__name__ not in _FUSABLE_MODELS:
        return None
    params = getattr(pc, "gaussians", None)
    if params is None or any(k not in params for k in _RAW_KEYS):
        return None
    raw = {k: params[k] for k in _RAW_KEYS}
    if raw["shs_rest"].shape[1] == 0 or raw["opacities"].dim() != 2:
        return None
    return raw


# ----------------------------------------------------------------------------------------------------------------------
# vanilla
# ----------------------------------------------------------------------------------------------------------------------
class B200VanillaRenderer(Renderer):
    def __init__(self, compute_cov3D_python: bool = False, convert_SHs_python: bool = False, cache_cameras: bool = True,
                 fused_activations: bool = True):
        """fused_activations: when the model is the vanilla Gaussian model (raw ``means / scales / rotations / opacities /
        shs_dc / shs_rest`` parameters with exp / normalize / sigmoid activations, vanilla_gaussian.py:345-358), feed the
        RAW parameters to the kernels and apply the activations there.  Any other model goes through its getters."""
        super().__init__()
        if compute_cov3D_python:
            raise NotImplementedError("b200gs computes cov3D in the projection kernel; compute_cov3D_python is unsupported")
        self.compute_cov3D_python = compute_cov3D_python
        self.convert_SHs_python = convert_SHs_python
        self.cache_cameras = cache_cameras
        self.fused_activations = fused_activations

    def forward(self, viewpoint_camera, pc, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                render_types: list = None, **kwargs):
        if render_types is None:
            render_types = ["rgb"]
        assert len(render_types) == 1, "Only single type is allowed currently"
        rendered_image_key = "render"
        if "depth" in render_types:
            rendered_image_key = "depth"
            w2c = viewpoint_camera.world_to_camera
            depth = (torch.matmul(pc.get_xyz, w2c[:3, :3]) + w2c[3, :3])[:, 2:]
            bg_color = torch.zeros_like(bg_color)
            override_color = depth.repeat(1, 3)
        raw = _raw_parameters(pc) if (self.fused_activations and override_color is None) else None
        if raw is not None:
            # fused path: exp / sigmoid / normalize / cat run inside K1 and K8 on the raw parameter tensors
            # a LEAF that requires grad: `.grad` is populated exactly as for the reference's `zeros_like(...) + 0` + retain_grad()
            # (vanilla_renderer.py:55-56; retain_grad() on a leaf is a no-op), without the extra elementwise kernel
            screenspace_points = torch.zeros_like(raw["means"], requires_grad=True)
            view = camera_view(viewpoint_camera, MODE_VANILLA, self.cache_cameras)
            view = _view_with(view, sh_degree=int(pc.active_sh_degree), scale_modifier=float(scaling_modifier))
            image, radii = ops.rasterize_vanilla_raw(raw["means"], screenspace_points, raw["shs_dc"], raw["shs_rest"],
                                                     raw["opacities"], raw["scales"], raw["rotations"], bg_color, view)
            return {
                rendered_image_key: image,
                "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0,
                "radii": radii,
            }
        out = self.render(pc.get_xyz, pc.get_opacity, pc.get_scaling, pc.get_rotation,
                          pc.get_features if override_color is None else None, pc.active_sh_degree, viewpoint_camera,
                          bg_color, scaling_modifier, colors_precomp=override_color, cache_cameras=self.cache_cameras)
        return {
            rendered_image_key: out["render"],
            "viewspace_points": out["viewspace_points"],
            "visibility_filter": out["visibility_filter"],
            "radii": out["radii"],
        }

    @staticmethod
    def render(means3D, opacity, scales, rotations, features, active_sh_degree: int, viewpoint_camera, bg_color: torch.Tensor,
               scaling_modifier=1.0, colors_precomp=None, cov3D_precomp=None, cache_cameras: bool = True):
        if colors_precomp is not None:
            assert features is None
        if cov3D_precomp is not None:
            raise NotImplementedError("b200gs: cov3D_precomp is unsupported")
        # zero tensor whose .grad receives dL/d(mean2D) (vanilla_renderer.py:55-56)
        screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=means3D.device) + 0
        view = camera_view(viewpoint_camera, MODE_VANILLA, cache_cameras)
        view = _view_with(view, sh_degree=int(active_sh_degree), scale_modifier=float(scaling_modifier))
        image, radii = ops.rasterize_vanilla(means3D, screenspace_points, features, colors_precomp, opacity, scales, rotations,
                                             bg_color, view)
        return {
            "render": image,
            "depth": None,
            "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0,
            "radii": radii,
        }

    def get_available_outputs(self) -> Dict:
        return {
            "rgb": RendererOutputInfo("render"),
            "depth": RendererOutputInfo("depth", RendererOutputTypes.GRAY),
        }


# ----------------------------------------------------------------------------------------------------------------------
# gsplat
# ----------------------------------------------------------------------------------------------------------------------
_GRAD_SCALES = {}


def _grad_scale(width: int, height: int, device) -> torch.Tensor:
    """0.5 * [[W, H]] on the device (gsplat_renderer.py:200), built once per (size, device): a fresh torch.tensor(...).to(device)
    every step is a pageable host->device copy, i.e. a stream synchronisation in the middle of the step."""
    key = (int(width), int(height), str(device))
    t = _GRAD_SCALES.get(key)
    if t is None:
        t = _GRAD_SCALES[key] = 0.5 * torch.tensor([[float(width), float(height)]], dtype=torch.float32).to(device)
    return t


class B200GSplatRenderer(Renderer):
    _RGB_REQUIRED = 1
    _ALPHA_REQUIRED = 1 << 1
    _ACC_DEPTH_REQUIRED = 1 << 2
    _ACC_DEPTH_INVERTED_REQUIRED = 1 << 3
    _EXP_DEPTH_REQUIRED = 1 << 4
    _EXP_DEPTH_INVERTED_REQUIRED = 1 << 5
    _INVERSE_DEPTH_REQUIRED = 1 << 6
    _HARD_DEPTH_REQUIRED = 1 << 7
    _HARD_INVERSE_DEPTH_REQUIRED = 1 << 8

    RENDER_TYPE_BITS = {
        "rgb": _RGB_REQUIRED,
        "alpha": _ALPHA_REQUIRED | _ACC_DEPTH_REQUIRED,
        "acc_depth": _ACC_DEPTH_REQUIRED,
        "acc_depth_inverted": _ACC_DEPTH_REQUIRED | _ACC_DEPTH_INVERTED_REQUIRED,
        "exp_depth": _ACC_DEPTH_REQUIRED | _EXP_DEPTH_REQUIRED,
        "exp_depth_inverted": _ACC_DEPTH_REQUIRED | _EXP_DEPTH_REQUIRED | _EXP_DEPTH_INVERTED_REQUIRED,
        "inverse_depth": _INVERSE_DEPTH_REQUIRED,
        "hard_depth": _HARD_DEPTH_REQUIRED,
        "hard_inverse_depth": _HARD_INVERSE_DEPTH_REQUIRED,
    }

    def __init__(self, block_size: int = DEFAULT_BLOCK_SIZE, anti_aliased: bool = DEFAULT_ANTI_ALIASED_STATUS,
                 kernel_size: float = 0.3, cache_cameras: bool = True, fused_activations: bool = True, absgrad: bool = False) -> None:
        """absgrad: also accumulate |dL/dmean2D| per Gaussian and hand it out as ``viewspace_points.absgrad`` after backward — what the
        density controller reads when its ``absgrad`` option is on (vanilla_density_controller.py:112-113); the rgb pass computes it
        (the op-by-op path: the fused single-node path has no absgrad output)."""
        super().__init__()
        if block_size != DEFAULT_BLOCK_SIZE:
            raise ValueError("b200gs supports block_size 16 only")
        self.absgrad = bool(absgrad)
        self.block_size = block_size
        self.anti_aliased = anti_aliased
        self.filter_2d_kernel_size = kernel_size
        self.cache_cameras = cache_cameras
        self.fused_activations = fused_activations   # rgb-only renders of a vanilla Gaussian model: activations + SH inside K1/K8

    def parse_render_types(self, render_types: list) -> int:
        if render_types is None:
            return self._RGB_REQUIRED
        bits = 0
        for i in render_types:
            bits |= self.RENDER_TYPE_BITS[i]
        return bits

    @staticmethod
    def is_type_required(bits: int, type: int) -> bool:
        return bits & type != 0

    @staticmethod
    def _project(means3D, scales, rotations, viewpoint_camera, scaling_modifier, eps2d=0.3, cache_cameras=True):
        view = camera_view(viewpoint_camera, MODE_GSPLAT, cache_cameras)
        view = _view_with(view, scale_modifier=float(scaling_modifier), eps2d=float(eps2d))
        return ops.project_gaussians(means3D, scales, scaling_modifier, rotations, None, 0, 0, 0, 0, view.height, view.width,
                                     view=view)

    def forward(self, viewpoint_camera, pc, bg_color: torch.Tensor, scaling_modifier=1.0, render_types: list = None, **kwargs):
        bits = self.parse_render_types(render_types)
        raw = _raw_parameters(pc) if (self.fused_activations and bits == self._RGB_REQUIRED and not getattr(self, "absgrad", False)) else None
        if raw is not None:
            # one autograd node for the whole step (K1..K6 / K7 -> gradient rows -> K8); image size from the cached host view: no
            # device->host read, no pageable host->device copy anywhere in the step
            view = camera_view(viewpoint_camera, MODE_GSPLAT, self.cache_cameras)
            view = _view_with(view, scale_modifier=float(scaling_modifier), eps2d=float(getattr(self, "filter_2d_kernel_size", 0.3)),
                              sh_degree=int(pc.active_sh_degree))
            img_height, img_width = int(view.height), int(view.width)
            xys = torch.zeros(raw["means"].shape[0], 2, dtype=torch.float32, device=raw["means"].device, requires_grad=True)
            image, radii = ops.rasterize_gsplat_raw(raw["means"], xys, raw["shs_dc"], raw["shs_rest"], raw["opacities"], raw["scales"],
                                                    raw["rotations"], bg_color, view, self.anti_aliased)
            rgb = image.permute(2, 0, 1)
            none = None
            return {
                "render": rgb, "alpha": none, "acc_depth": none, "acc_depth_inverted": none, "exp_depth": none,
                "exp_depth_inverted": none, "inverse_depth": none, "hard_depth": none, "hard_inverse_depth": none,
                "viewspace_points": xys,
                "viewspace_points_grad_scale": _grad_scale(img_width, img_height, xys.device),
                "visibility_filter": radii > 0,
                "radii": radii,
            }
        img_height, img_width = int(viewpoint_camera.height), int(viewpoint_camera.width)
        quats = pc.get_rotation
        quats = quats / quats.norm(dim=-1, keepdim=True)  # gsplat_renderer.py:68
        xys, depths, radii, conics, comp, num_tiles_hit, cov3d = self._project(
            pc.get_xyz, pc.get_scaling, quats, viewpoint_camera, scaling_modifier,
            getattr(self, "filter_2d_kernel_size", 0.3), self.cache_cameras)

        opacities = pc.get_opacity
        if self.anti_aliased is True:
            opacities = opacities * comp[:, None]

        def rasterize(feats, background, return_alpha=False, opac=opacities):
            return ops.rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, feats, opac, img_height, img_width,
                                           self.block_size, background, return_alpha)

        rgb = None
        if self.is_type_required(bits, self._RGB_REQUIRED):
            viewdirs = pc.get_xyz.detach() - viewpoint_camera.camera_center
            rgbs = ops.spherical_harmonics(pc.active_sh_degree, viewdirs, pc.get_features)
            rgbs = torch.clamp(rgbs + 0.5, min=0.0)
            if getattr(self, "absgrad", False):
                rgb = ops.rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, rgbs, opacities, img_height, img_width, self.block_size,
                                              bg_color, False, absgrad=True).permute(2, 0, 1)
            else:
                rgb = rasterize(rgbs, bg_color).permute(2, 0, 1)

        alpha = acc_depth_im = acc_depth_inverted_im = exp_depth_im = exp_depth_inverted_im = None
        zero1 = torch.zeros((1,), dtype=torch.float, device=bg_color.device)
        if self.is_type_required(bits, self._ACC_DEPTH_REQUIRED):
            acc_depth_im, alpha = rasterize(depths.unsqueeze(-1), zero1, True)
            alpha = alpha[..., None]
            if self.is_type_required(bits, self._ACC_DEPTH_INVERTED_REQUIRED):
                acc_depth_inverted_im = torch.where(acc_depth_im > 0, 1. / acc_depth_im, acc_depth_im.detach().max()).permute(2, 0, 1)
            if self.is_type_required(bits, self._EXP_DEPTH_REQUIRED):
                exp_depth_im = torch.where(alpha > 0, acc_depth_im / alpha, acc_depth_im.detach().max()).permute(2, 0, 1)
            alpha = alpha.permute(2, 0, 1) if self.is_type_required(bits, self._ALPHA_REQUIRED) else None
            acc_depth_im = acc_depth_im.permute(2, 0, 1)
            if self.is_type_required(bits, self._EXP_DEPTH_INVERTED_REQUIRED):
                exp_depth_inverted_im = torch.where(exp_depth_im > 0, 1. / exp_depth_im, exp_depth_im.detach().max())

        inverse_depth_im = None
        if self.is_type_required(bits, self._INVERSE_DEPTH_REQUIRED):
            inverse_depth = 1. / (depths.clamp_min(0.) + 1e-8).unsqueeze(-1)
            inverse_depth_im = rasterize(inverse_depth, zero1).permute(2, 0, 1)

        hard_depth_im = None
        if self.is_type_required(bits, self._HARD_DEPTH_REQUIRED):
            hard_depth_im = rasterize(depths.unsqueeze(-1), zero1, False, opacities + (1 - opacities.detach())).permute(2, 0, 1)

        hard_inverse_depth_im = None
        if self.is_type_required(bits, self._HARD_INVERSE_DEPTH_REQUIRED):
            inverse_depth = 1. / (depths.clamp_min(0.) + 1e-8).unsqueeze(-1)
            hard_inverse_depth_im = rasterize(inverse_depth, zero1, False, opacities + (1 - opacities.detach())).permute(2, 0, 1)

        return {
            "render": rgb,
            "alpha": alpha,
            "acc_depth": acc_depth_im,
            "acc_depth_inverted": acc_depth_inverted_im,
            "exp_depth": exp_depth_im,
            "exp_depth_inverted": exp_depth_inverted_im,
            "inverse_depth": inverse_depth_im,
            "hard_depth": hard_depth_im,
            "hard_inverse_depth": hard_inverse_depth_im,
            "viewspace_points": xys,
            "viewspace_points_grad_scale": _grad_scale(img_width, img_height, xys.device),
            "visibility_filter": radii > 0,
            "radii": radii,
        }

    @staticmethod
    def render(means3D, opacities, scales, rotations, features, active_sh_degree: int, viewpoint_camera, bg_color: torch.Tensor,
               scaling_modifier=1.0, anti_aliased: bool = DEFAULT_ANTI_ALIASED_STATUS, colors_precomp=None,
               color_computer=None, block_size: int = DEFAULT_BLOCK_SIZE, extra_projection_kwargs: dict = None):
        img_height, img_width = int(viewpoint_camera.height), int(viewpoint_camera.width)
        eps2d = (extra_projection_kwargs or {}).get("filter_2d_kernel_size", 0.3)
        xys, depths, radii, conics, comp, num_tiles_hit, cov3d = B200GSplatRenderer._project(
            means3D, scales, rotations, viewpoint_camera, scaling_modifier, eps2d)
        if colors_precomp is not None:
            rgbs = colors_precomp
        elif color_computer is not None:
            rgbs = color_computer(locals())
        else:
            viewdirs = means3D.detach() - viewpoint_camera.camera_center
            rgbs = torch.clamp(ops.spherical_harmonics(active_sh_degree, viewdirs, features) + 0.5, min=0.0)
        if anti_aliased is True:
            opacities = opacities * comp[:, None]
        rgb = ops.rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, rgbs, opacities, img_height, img_width,
                                      block_size, bg_color, False)
        return {
            "render": rgb.permute(2, 0, 1),
            "viewspace_points": xys,
            "viewspace_points_grad_scale": _grad_scale(img_width, img_height, xys.device),
            "visibility_filter": radii > 0,
            "radii": radii,
        }

    @staticmethod
    def project(means3D, scales, rotations, viewpoint_camera, scaling_modifier=1.0, block_size: int = DEFAULT_BLOCK_SIZE,
                extra_projection_kwargs: dict = None):
        eps2d = (extra_projection_kwargs or {}).get("filter_2d_kernel_size", 0.3)
        return B200GSplatRenderer._project(means3D, scales, rotations, viewpoint_camera, scaling_modifier, eps2d)

    @staticmethod
    def rasterize_simplified(project_results, viewpoint_camera, colors, bg_color, opacities, anti_aliased: bool = True):
        xys, depths, radii, conics, comp, num_tiles_hit, cov3d = project_results
        if anti_aliased is True:
            opacities = opacities * comp[:, None]
        return ops.rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacities, int(viewpoint_camera.height),
                                       int(viewpoint_camera.width), DEFAULT_BLOCK_SIZE, bg_color, False).permute(2, 0, 1)

    @staticmethod
    def rasterize(opacities, rgbs, bg_color, project_results: Tuple, viewpoint_camera, xys_retain_grad: bool = True,
                  block_size: int = DEFAULT_BLOCK_SIZE, anti_aliased: bool = DEFAULT_ANTI_ALIASED_STATUS):
        img_height, img_width = int(viewpoint_camera.height), int(viewpoint_camera.width)
        xys, depths, radii, conics, comp, num_tiles_hit, cov3d = project_results
        if xys_retain_grad is True:
            try:
                xys.retain_grad()
            except Exception:
                pass
        if anti_aliased is True:
            opacities = opacities * comp[:, None]
        rgb = ops.rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, rgbs, opacities, img_height, img_width,
                                      block_size, bg_color, False)
        return {
            "render": rgb.permute(2, 0, 1),
            "viewspace_points": xys,
            "viewspace_points_grad_scale": _grad_scale(img_width, img_height, xys.device),
            "visibility_filter": radii > 0,
            "radii": radii,
        }

    def get_available_outputs(self) -> Dict:
        gray = RendererOutputTypes.GRAY
        return {
            "rgb": RendererOutputInfo("render"),
            "alpha": RendererOutputInfo("alpha", type=gray),
            "acc_depth": RendererOutputInfo("acc_depth", type=gray),
            "acc_depth_inverted": RendererOutputInfo("acc_depth_inverted", type=gray),
            "exp_depth": RendererOutputInfo("exp_depth", type=gray),
            "exp_depth_inverted": RendererOutputInfo("exp_depth_inverted", type=gray),
            "inverse_depth": RendererOutputInfo("inverse_depth", type=gray),
            "hard_depth": RendererOutputInfo("hard_depth", type=gray),
            "hard_inverse_depth": RendererOutputInfo("hard_inverse_depth", type=gray),
        }


@dataclass
class B200VanillaRendererConfig(RendererConfig):
    compute_cov3D_python: bool = False
    convert_SHs_python: bool = False

    def instantiate(self, *args, **kwargs) -> Renderer:
        return B200VanillaRenderer(self.compute_cov3D_python, self.convert_SHs_python)


@dataclass
class B200GSplatRendererConfig(RendererConfig):
    block_size: int = DEFAULT_BLOCK_SIZE
    anti_aliased: bool = DEFAULT_ANTI_ALIASED_STATUS
    kernel_size: float = 0.3
    absgrad: bool = False

    def instantiate(self, *args, **kwargs) -> Renderer:
        return B200GSplatRenderer(self.block_size, self.anti_aliased, self.kernel_size, absgrad=self.absgrad)
