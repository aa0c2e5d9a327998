"""b200gs — a Blackwell (sm_100a) differentiable 3D-Gaussian rasterizer behind gaussian-splatting-lightning's
``internal/renderers`` plug-in surface.

Layout:
  csrc/        CUDA kernels + the C ABI (``libb200gs.so``; declared in ``include/b200gs.h``)
  _lib.py      ctypes binding of the C ABI (fails loudly if the library is missing)
  ops.py       torch.autograd.Function wrappers (device memory / streams only; no math in torch)
  compat/      drop-in modules with the names the reference imports (diff_gaussian_rasterization, gsplat v0 API)
  renderers.py Renderer plug-ins mirroring VanillaRenderer / GSPlatRenderer
  distributed.py  Gaussian-sharded multi-GPU renderer (all-to-all of projected splats)
  cameras.py, scene.py  camera container + the deterministic synthetic scene generator used by tests and bench
"""
__version__ = "0.1.0"
