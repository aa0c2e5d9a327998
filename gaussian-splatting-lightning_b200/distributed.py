"""Gaussian-sharded multi-GPU renderer — the reference's exchange, on the b200gs kernels.

Mirrors ``internal/renderers/gsplat_distributed_renderer.py``:

* parameters are sharded by contiguous Gaussian index ranges (``:76-89``); there is no gradient all-reduce;
* every rank renders ONE camera per step; the W camera descriptions are all-gathered (``:319-335`` gathers ids and looks
  them up in the dataset — here the 40-float view itself is gathered, so no dataset object is needed);
* each rank projects its shard to all W cameras (K1, gsplat constants) and evaluates SH colours (``:252-311``);
* the VISIBLE projected splats — 11 floats (xy 2, depth 1, conic 3, compensation 1, opacity 1, rgb 3) + radius — are sent
  to the rank that owns the camera with an all-to-all (``:127-217``).  Here: ONE ``all_to_all_single`` of ``[V, 12]``
  fp32 rows (radius bit-cast into the 12th column) instead of the reference's two list-form all-to-alls — one message
  per peer, and it also runs on gloo for the CPU tests (gloo has no list-form all_to_all);
* the owner concatenates the rows in rank order (= global Gaussian-index order, so depth ties break exactly as in the
  single-GPU renderer), bins and blends its whole image locally (K2-K7);
* backward is the mirror image: blend backward on the owner, the ``[V,12]`` gradient rows travel back through the same
  all-to-all, K8 runs per camera on the shard's owner.

The distributed image is bit-identical to the single-GPU gsplat-mode render of the unsharded model
(tests/test_gpu_distributed.py).  A per-pixel reduce of partial images would NOT be.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

ROW_FLOATS = 12  # xy(2) depth(1) conic(3) comp(1) opacity(1) rgb(3) radius-bits(1)
VIEW_FLOATS = 40  # width height fx fy cx cy | world_to_camera 16 | camera_center 3 | pad


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous index shard of rank (gsplat_distributed_renderer.py:76-83): round(n/W) each, last takes the rest."""
    per = round(n / world_size)
    lo = per * rank
    hi = n if rank + 1 == world_size else lo + per
    return lo, min(hi, n)


class _AllToAllRows(torch.autograd.Function):
    """all_to_all_single of row blocks with uneven splits; the backward is the mirrored exchange."""

    @staticmethod
    def forward(ctx, rows: torch.Tensor, send_counts: List[int], recv_counts: List[int], group):
        rows = rows.contiguous()
        out = rows.new_empty((sum(recv_counts),) + tuple(rows.shape[1:]))
        dist.all_to_all_single(out, rows, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=group)
        ctx.send_counts, ctx.recv_counts, ctx.group = list(send_counts), list(recv_counts), group
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        g = g.contiguous()
        gin = g.new_empty((sum(ctx.send_counts),) + tuple(g.shape[1:]))
        dist.all_to_all_single(gin, g, output_split_sizes=ctx.send_counts, input_split_sizes=ctx.recv_counts, group=ctx.group)
        return gin, None, None, None


def exchange_counts(send_counts: Sequence[int], device, group=None) -> List[int]:
    """Tell every peer how many rows it will receive from me; returns how many I receive from each peer."""
    send = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return [int(x) for x in recv.tolist()]


def exchange_rows(rows_per_dest: Sequence[torch.Tensor], group=None) -> Tuple[torch.Tensor, List[int]]:
    """rows_per_dest[j]: [V_j, C] rows destined to rank j.  Returns (rows received, concatenated in source-rank order;
    per-source counts).  Differentiable w.r.t. the rows."""
    send_counts = [int(r.shape[0]) for r in rows_per_dest]
    recv_counts = exchange_counts(send_counts, rows_per_dest[0].device, group)
    rows = torch.cat(list(rows_per_dest), dim=0)
    return _AllToAllRows.apply(rows, send_counts, recv_counts, group), recv_counts


def pack_rows(xys, depths, conics, comp, opacities, rgbs, radii, visible) -> torch.Tensor:
    """[V, 12] fp32 rows of the visible splats (gsplat_distributed_renderer.py:167-178 + the int tensor of :177)."""
    radii_bits = radii.view(torch.float32) if radii.dtype == torch.int32 else radii.to(torch.int32).view(torch.float32)
    rows = torch.cat([xys, depths.unsqueeze(-1), conics, comp.unsqueeze(-1), opacities.reshape(-1, 1), rgbs,
                      radii_bits.unsqueeze(-1)], dim=-1)
    return rows[visible]


def unpack_rows(rows: torch.Tensor):
    xys, depths, conics, comp, opac, rgbs, rbits = torch.split(rows, [2, 1, 3, 1, 1, 3, 1], dim=-1)
    radii = rbits.detach().contiguous().view(torch.int32).squeeze(-1)
    return xys.contiguous(), depths.squeeze(-1).contiguous(), conics.contiguous(), comp.squeeze(-1), opac.squeeze(-1), rgbs.contiguous(), radii


def pack_view(camera) -> torch.Tensor:
    """The quantities K1 (gsplat constants) needs from a camera, as VIEW_FLOATS fp32 on the camera's device."""
    dev = camera.world_to_camera.device
    v = torch.zeros(VIEW_FLOATS, dtype=torch.float32, device=dev)
    head = torch.stack([camera.width.float(), camera.height.float(), camera.fx.float(), camera.fy.float(), camera.cx.float(),
                        camera.cy.float()]).to(dev)
    v[0:6] = head
    v[6:22] = camera.world_to_camera.reshape(-1)
    v[22:25] = camera.camera_center
    return v


class GatheredView:
    """Host-side view of one gathered camera (what ``cameras`` in the return dict holds)."""

    def __init__(self, flat: torch.Tensor, device):
        f = flat.tolist()
        self.width, self.height = int(f[0]), int(f[1])
        self.fx, self.fy, self.cx, self.cy = f[2], f[3], f[4], f[5]
        self.world_to_camera = flat[6:22].reshape(4, 4)
        self.camera_center = flat[22:25].to(device)


class B200DistributedRenderer(torch.nn.Module):
    """Drop-in for ``GSplatDistributedRendererImpl.forward`` (gsplat_distributed_renderer.py:313-414): `pc` holds THIS rank's
    shard; returns this rank's image plus the per-camera projection results the distributed density controller reads
    (``distributed_vanilla_density_controller.py:16-47``)."""

    def __init__(self, anti_aliased: bool = True, group=None):
        super().__init__()
        self.anti_aliased = anti_aliased
        self.group = group

    def forward(self, viewpoint_camera, pc, bg_color: torch.Tensor, scaling_modifier=1.0, render_types: list = None, **kwargs):
        from . import ops
        from ._lib import MODE_GSPLAT
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        dev = bg_color.device

        # 1. every rank learns all W cameras
        mine = pack_view(viewpoint_camera)
        gathered = torch.empty(world * VIEW_FLOATS, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        flat = gathered.reshape(world, VIEW_FLOATS).cpu()
        views = [GatheredView(flat[j], dev) for j in range(world)]

        # 2. project my shard to every camera, colours for every camera
        means, scales, opacities = pc.get_xyz, pc.get_scaling, pc.get_opacity
        quats = pc.get_rotation
        quats = quats / quats.norm(dim=-1, keepdim=True)
        feats = pc.get_features
        rows_per_dest, projection_results_list, visible_mask_list = [], [], []
        for j, gv in enumerate(views):
            view = ops.make_view(MODE_GSPLAT, gv.width, gv.height, fx=gv.fx, fy=gv.fy, cx=gv.cx, cy=gv.cy,
                                 viewmatrix=gv.world_to_camera, scale_modifier=scaling_modifier)
            xys, depths, radii, conics, comp, tiles, _ = ops.project_gaussians(means, scales, scaling_modifier, quats, None, 0, 0, 0, 0,
                                                                              gv.height, gv.width, view=view)
            visible = radii > 0
            rgbs = torch.clamp(ops.spherical_harmonics(pc.active_sh_degree, means.detach() - gv.camera_center, feats) + 0.5, min=0.0)
            rows_per_dest.append(pack_rows(xys, depths, conics, comp, opacities, rgbs, radii, visible))
            projection_results_list.append((radii, xys, depths, conics, comp, visible))
            visible_mask_list.append(visible)

        # 3. all-to-all of the visible splats
        rows, recv_counts = exchange_rows(rows_per_dest, self.group)

        # 4. local rasterization of my camera
        xys, depths, conics, comp, opac, rgbs, radii = unpack_rows(rows)
        if self.anti_aliased:
            opac = opac * comp
        gv = views[rank]
        img = ops.rasterize_gaussians(xys, depths, radii, conics, None, rgbs, opac, gv.height, gv.width, 16, bg_color, False)
        return {
            "render": img.permute(2, 0, 1),
            "cameras": views,
            "projection_results_list": projection_results_list,
            "visible_mask_list": visible_mask_list,
            "xys_grad_scale_required": True,
            "n_received": recv_counts,
        }
