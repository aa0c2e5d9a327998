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
