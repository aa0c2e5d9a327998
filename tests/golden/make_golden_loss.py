"""Freeze outputs of the reference's OWN loss functions (internal/utils/ssim.py, imported by file path from /root/reference in the
authoring container — the GPU box has no /root/reference) for seeded image pairs: SSIM, L1, the combined training loss of
vanilla_metrics.py:57-74 (lambda_dssim 0.2) and its autograd gradient w.r.t. the rendered image.
Run:  python tests/golden/make_golden_loss.py      -> tests/golden/loss_*.npz"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_ssim", "/root/reference/internal/utils/ssim.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def pair(seed, H, W):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1)      # a plausible render: the target plus noise
    return img, gt


for seed, H, W in [(0, 37, 50), (1, 96, 128), (2, 11, 11), (3, 5, 200)]:
    img, gt = pair(seed, H, W)
    img.requires_grad_(True)
    l1 = ref.l1_loss(img, gt)
    ss = ref.ssim(img, gt)
    loss = 0.8 * l1 + 0.2 * (1.0 - ss)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, f"loss_s{seed}_{H}x{W}.npz"), meta=np.array([seed, H, W]), l1=l1.detach().numpy(),
                        ssim=ss.detach().numpy(), loss=loss.detach().numpy(), grad=img.grad.numpy())
    print(seed, H, W, float(l1), float(ss), float(loss))
