"""Golden vectors for SH degree 4 from the reference's own sh_utils.eval_sh (internal/utils/sh_utils.py:57-112), run in the
authoring container:  python tests/golden/make_golden_sh4.py   -> tests/golden/sh_deg4.npz (values + autograd gradients)."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_sh_utils", "/root/reference/internal/utils/sh_utils.py")
sh = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sh)

g = torch.Generator().manual_seed(44)
n = 64
shs = (0.3 * torch.randn(n, 25, 3, generator=g)).requires_grad_(True)
dirs = torch.randn(n, 3, generator=g)
dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).requires_grad_(True)
cot = torch.randn(n, 3, generator=g)
rgb = sh.eval_sh(4, shs.transpose(1, 2), dirs)
rgb2 = sh.eval_sh_decomposed(4, shs[:, :1, :], shs[:, 1:, :], dirs)
assert torch.allclose(rgb, rgb2)      # the identity the reference's own tests/sh_utils_test.py pins
(rgb * cot).sum().backward()
np.savez(os.path.join(HERE, "sh_deg4.npz"), shs=shs.detach().numpy(), dirs=dirs.detach().numpy(), cot=cot.numpy(), rgb=rgb.detach().numpy(),
         g_shs=shs.grad.numpy(), g_dirs=dirs.grad.numpy())
print("wrote sh_deg4.npz", float(rgb.abs().max()))
