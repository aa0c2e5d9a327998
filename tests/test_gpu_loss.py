"""Fused L1 + SSIM training loss (csrc/loss.cu) against the CPU oracle (oracle/loss_oracle.py, pinned to the
reference's internal/utils/ssim.py + vanilla_metrics.py:57-74) and against golden vectors produced by the reference itself
(tests/golden/make_golden_loss.py)."""
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(37, 50), (96, 128), (11, 11), (5, 200), (270, 480)])
def test_fused_l1_ssim_loss_matches_oracle(H, W):
    import torch
    from b200gs import ops
    from oracle import loss_oracle as LO
    g = torch.Generator().manual_seed(H * 1000 + W)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    ref_img = img.double().requires_grad_(True)
    ref_loss, ref_l1, ref_ssim = LO.training_loss(ref_img, gt.double())
    (3.0 * ref_loss).backward()
    dimg = img.cuda().requires_grad_(True)
    loss, stats = ops.l1_ssim_loss(dimg, gt.cuda(), 0.2)
    (3.0 * loss).backward()
    assert abs(float(loss) - float(ref_loss)) < 2e-6
    assert abs(float(stats[0]) - float(ref_l1)) < 1e-6 and abs(float(stats[1]) - float(ref_ssim)) < 5e-6
    err = (dimg.grad.double().cpu() - ref_img.grad).abs().max() / ref_img.grad.abs().max()
    assert float(err) < 1e-3, float(err)


def test_fused_l1_ssim_loss_matches_reference_golden():
    import glob

    import numpy as np
    import torch
    from b200gs import ops
    from conftest import GOLDEN
    for path in sorted(glob.glob(os.path.join(GOLDEN, "loss_*.npz"))):
        d = np.load(path)
        seed, H, W = (int(x) for x in d["meta"])
        g = torch.Generator().manual_seed(seed)
        gt = torch.rand(3, H, W, generator=g)
        img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1).cuda().requires_grad_(True)
        loss, stats = ops.l1_ssim_loss(img, gt.cuda(), 0.2)
        loss.backward()
        assert abs(float(loss) - float(d["loss"])) < 3e-6 and abs(float(stats[1]) - float(d["ssim"])) < 1e-5
        gref = torch.tensor(d["grad"])
        assert float((img.grad.cpu() - gref).abs().max() / gref.abs().max()) < 1e-3
