"""CPU restatement of the training loss that follows the rasterizer every step — TEST INFRASTRUCTURE ONLY (SURVEY §8f row 2:
the fused L1 + SSIM kernel is the next row of the hot-path table; this oracle and its golden vectors come first).

Follows the reference:
  * `internal/utils/ssim.py:17-18`   l1_loss = mean |a - b|
  * `internal/utils/ssim.py:23-31`   11-tap Gaussian window, sigma 1.5, normalised; the 2-D window is its outer product
  * `internal/utils/ssim.py:43-63`   SSIM map from five zero-padded depthwise convolutions, C1 = 0.01^2, C2 = 0.03^2, mean
  * `internal/metrics/vanilla_metrics.py:57-74`  loss = (1 - lambda) * L1 + lambda * (1 - SSIM), lambda_dssim = 0.2

The window is applied SEPARABLY here (rows then columns) — the formulation a fused CUDA kernel uses — which equals the reference's
2-D convolution up to fp rounding; tests/test_loss_oracle_golden.py pins it against outputs of the reference's own functions.
dtype-generic: float32 to mimic the reference, float64 as the mathematical truth.  Differentiable through torch autograd.
"""
import math

import torch
import torch.nn.functional as F

WINDOW = 11
SIGMA = 1.5
C1 = 0.01 ** 2
C2 = 0.03 ** 2
LAMBDA_DSSIM = 0.2


def window_1d(dtype=torch.float32) -> torch.Tensor:
    g = torch.tensor([math.exp(-(x - WINDOW // 2) ** 2 / float(2 * SIGMA ** 2)) for x in range(WINDOW)], dtype=torch.float32)
    return (g / g.sum()).to(dtype)      # the reference builds it in float32 and casts (`ssim.py:23-25,39`)


def _blur(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Zero-padded separable 11x11 Gaussian blur of x [C,H,W]."""
    c = x.shape[0]
    x = x.unsqueeze(0)
    kh = w.reshape(1, 1, 1, WINDOW).expand(c, 1, 1, WINDOW)
    kv = w.reshape(1, 1, WINDOW, 1).expand(c, 1, WINDOW, 1)
    x = F.conv2d(x, kh, padding=(0, WINDOW // 2), groups=c)
    x = F.conv2d(x, kv, padding=(WINDOW // 2, 0), groups=c)
    return x.squeeze(0)


def ssim_map(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    w = window_1d(img1.dtype)
    mu1, mu2 = _blur(img1, w), _blur(img2, w)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(img1 * img1, w) - mu1_sq
    s2 = _blur(img2 * img2, w) - mu2_sq
    s12 = _blur(img1 * img2, w) - mu12
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    return ssim_map(img1, img2).mean()


def l1(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    return (img1 - img2).abs().mean()


def training_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = LAMBDA_DSSIM):
    """Returns (loss, l1, ssim) exactly as `VanillaMetricsImpl._get_basic_metrics` combines them."""
    a, s = l1(image, gt), ssim(image, gt)
    return (1.0 - lambda_dssim) * a + lambda_dssim * (1.0 - s), a, s
