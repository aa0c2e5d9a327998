"""oracle/loss_oracle.py (separable L1 + SSIM, the formulation of the planned fused kernel — SURVEY §8f row 2) against the
outputs of the reference's own `internal/utils/ssim.py` + the loss combination of `vanilla_metrics.py:57-74`, frozen by
tests/golden/make_golden_loss.py: values and the autograd gradient w.r.t. the rendered image."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import loss_oracle as L

FILES = sorted(glob.glob(os.path.join(GOLDEN, "loss_*.npz")))


def _pair(seed, H, W):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    return img, gt


def test_golden_files_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_loss_oracle_matches_reference(path, dtype):
    d = np.load(path)
    seed, H, W = (int(x) for x in d["meta"])
    img, gt = _pair(seed, H, W)
    img = img.to(dtype).requires_grad_(True)
    loss, l1, ssim = L.training_loss(img, gt.to(dtype))
    loss.backward()
    # the reference evaluates a 2-D 11x11 convolution in fp32; the separable form differs by rounding only
    assert abs(float(l1.detach()) - float(d["l1"])) < 2e-7
    assert abs(float(ssim.detach()) - float(d["ssim"])) < 5e-6
    assert abs(float(loss.detach()) - float(d["loss"])) < 2e-6
    g_ref = torch.tensor(d["grad"], dtype=torch.float64)
    err = (img.grad.double() - g_ref).abs().max() / g_ref.abs().max()
    assert float(err) < 2e-4, float(err)


def test_ssim_properties():
    img, gt = _pair(5, 40, 56)
    assert abs(float(L.ssim(gt, gt)) - 1.0) < 1e-6                       # identical images
    assert abs(float(L.ssim(img, gt)) - float(L.ssim(gt, img))) < 1e-7   # symmetric
    w = L.window_1d()
    assert abs(float(w.sum()) - 1.0) < 1e-6 and w.numel() == 11 and float(w[5]) == float(w.max())


def test_fused_kernel_formulas_match_autograd():
    """The closed-form pieces the fused kernels (csrc/loss.cu) evaluate — the three partial-derivative maps of the SSIM map
    w.r.t. the blurred moments mu1, E[a^2], E[ab], and  v_img = ((1-l) sign(a-b) - l (blur(d_mu1) + 2 a blur(d_e11) + b blur(d_e12))) / n
    — restated with torch ops and compared with autograd through the oracle loss.  Pins the math of K9/K10 on CPU."""
    g = torch.Generator().manual_seed(7)
    gt = torch.rand(3, 45, 61, generator=g, dtype=torch.float64)
    a = (gt + 0.15 * torch.randn(3, 45, 61, generator=g, dtype=torch.float64)).clamp(0, 1)
    lam = 0.2
    ar = a.clone().requires_grad_(True)
    loss, _, _ = L.training_loss(ar, gt, lam)
    loss.backward()

    w = L.window_1d(torch.float64)
    mu1, mu2 = L._blur(a, w), L._blur(gt, w)
    e11, e22, e12 = L._blur(a * a, w), L._blur(gt * gt, w), L._blur(a * gt, w)
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + L.C1, 2 * s12 + L.C2
    B1, B2 = mu1 * mu1 + mu2 * mu2 + L.C1, s1 + s2 + L.C2
    ss = A1 * A2 / (B1 * B2)
    d_mu1 = (2 * mu2 * A2 - 2 * mu2 * A1) / (B1 * B2) - ss * (2 * mu1 / B1 - 2 * mu1 / B2)
    d_e11 = -ss / B2
    d_e12 = 2 * A1 / (B1 * B2)
    n = a.numel()
    v_img = ((1 - lam) * torch.sign(a - gt) - lam * (L._blur(d_mu1, w) + 2 * a * L._blur(d_e11, w) + gt * L._blur(d_e12, w))) / n
    assert float((v_img - ar.grad).abs().max() / ar.grad.abs().max()) < 1e-9
