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


class _ShardedRasterize(torch.autograd.Function):
    """The whole sharded step as ONE autograd node, everything between the raw shard parameters and this rank's image:
    K1 (fused activations, gsplat constants) per camera -> pack visible rows (device-side compaction) -> count exchange
    (the step's single host sync) -> all_to_all_single of [V,12] rows -> K2-K7 reading the received rows in place.
    Backward: K7 accumulates straight into a [R,12] gradient row buffer -> mirrored all_to_all_single -> unpack -> K8
    (fused activation chain) per camera, summed over cameras."""

    @staticmethod
    def forward(ctx, means, log_scales, raw_quats, opac_logits, shs_dc, shs_rest, bg, views, rank, group, anti_aliased, sh_degree):
        import ctypes
        from . import ops
        from ._lib import MODE_GSPLAT, check, lib, ptr
        L = lib()
        dev = means.device
        n = means.shape[0]
        world = len(views)
        st = torch.cuda.current_stream().cuda_stream
        means, log_scales, raw_quats = means.contiguous(), log_scales.contiguous(), raw_quats.contiguous()
        ol = opac_logits.contiguous().reshape(-1)
        shs_dc, shs_rest, bg = shs_dc.contiguous(), shs_rest.contiguous(), bg.contiguous()
        counts = torch.zeros(2 * world, dtype=torch.int64, device=dev)   # [send | recv]
        ws = torch.empty(max(int(L.b200gs_pack_rows_workspace_bytes(n)), 256), dtype=torch.uint8, device=dev)
        per_cam = []
        for j, view in enumerate(views):
            v = ops._copy_view(view, sh_degree=int(sh_degree), sh_stride=int(shs_dc.shape[1] + shs_rest.shape[1]))
            xy, depth, radii, conic, comp, tiles, rgb, clamped, opac = ops.project_forward_raw(
                v, means, log_scales, raw_quats, ol, shs_dc, shs_rest, anti_aliased, want_comp=True)
            rows = torch.empty(n, ROW_FLOATS, dtype=torch.float32, device=dev)
            offsets = torch.empty(n, dtype=torch.int32, device=dev)
            check(L.b200gs_pack_rows(n, ptr(xy), ptr(depth), ptr(conic), ptr(comp), ptr(opac), ptr(rgb), ptr(radii), ptr(ws), ws.numel(),
                                     ptr(offsets), ptr(rows), counts.data_ptr() + 8 * j, st), "b200gs_pack_rows")
            per_cam.append((v, rows, offsets, radii, clamped, xy))
        dist.all_to_all_single(counts[world:], counts[:world], group=group)
        host_counts = counts.cpu().tolist()                                 # the step's host sync
        send_counts, recv_counts = host_counts[:world], host_counts[world:]
        send = torch.cat([per_cam[j][1][:send_counts[j]] for j in range(world)], dim=0)
        recv = torch.empty(sum(recv_counts), ROW_FLOATS, dtype=torch.float32, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
        del send

        gv = views[rank]
        W, H = gv.width, gv.height
        R = recv.shape[0]
        gx, gy = (W + 15) // 16, (H + 15) // 16
        ws_a = torch.empty(L.b200gs_bin_count_workspace_bytes(R), dtype=torch.uint8, device=dev)
        d_total = torch.empty(1, dtype=torch.int64, device=dev)
        host_total = torch.zeros(1, dtype=torch.int64).pin_memory()
        with ops._stage("bin_count"):
            check(L.b200gs_bin_count_rows(MODE_GSPLAT, W, H, R, ptr(recv), 1, ptr(ws_a), ws_a.numel(), ptr(d_total), host_total.data_ptr(), 1, st),
                  "b200gs_bin_count_rows")
        total = int(host_total[0])
        sorted_ids = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        ranges = torch.empty(gx * gy, 2, dtype=torch.int32, device=dev)
        ws_b = torch.empty(L.b200gs_bin_sort_workspace_bytes(R, total, W, H), dtype=torch.uint8, device=dev)
        with ops._stage("bin_sort"):
            check(L.b200gs_bin_sort_rows(MODE_GSPLAT, W, H, R, ptr(recv), 1, total, ptr(d_total), total, ptr(ws_a), ptr(ws_b), ws_b.numel(),
                                         ptr(sorted_ids), ptr(ranges), st), "b200gs_bin_sort_rows")
        image = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
        final_T = torch.empty(H, W, dtype=torch.float32, device=dev)
        n_contrib = torch.empty(H, W, dtype=torch.int32, device=dev)
        with ops._stage("blend_fwd"):
            check(L.b200gs_blend_fwd_rows(MODE_GSPLAT, W, H, ptr(ranges), ptr(sorted_ids), ptr(recv), ptr(bg), ptr(image), 3, 1, ptr(final_T),
                                          ptr(n_contrib), None, st), "b200gs_blend_fwd_rows")
        ctx.per_cam, ctx.views, ctx.rank, ctx.group = per_cam, views, rank, group
        ctx.counts = (send_counts, recv_counts)
        ctx.aa = bool(anti_aliased)
        ctx.hw = (H, W)
        ctx.opac_shape = tuple(opac_logits.shape)
        ctx.save_for_backward(means, log_scales, raw_quats, ol, shs_dc, shs_rest, bg, recv, sorted_ids, ranges, final_T, n_contrib)
        ctx.xy_grads = None
        return image

    @staticmethod
    def backward(ctx, v_image):
        from ._lib import MODE_GSPLAT, check, lib, ptr
        from . import ops
        L = lib()
        means, log_scales, raw_quats, ol, shs_dc, shs_rest, bg, recv, sorted_ids, ranges, final_T, n_contrib = ctx.saved_tensors
        dev = means.device
        n = means.shape[0]
        st = torch.cuda.current_stream().cuda_stream
        H, W = ctx.hw
        send_counts, recv_counts = ctx.counts
        v_image = v_image.contiguous()
        v_recv = torch.zeros_like(recv)
        with ops._stage("blend_bwd"):
            check(L.b200gs_blend_bwd_rows(MODE_GSPLAT, W, H, ptr(ranges), ptr(sorted_ids), ptr(recv), ptr(bg), ptr(final_T), ptr(n_contrib),
                                          ptr(v_image), 3, 1, None, ptr(v_recv), st), "b200gs_blend_bwd_rows")
        v_send = torch.empty(sum(send_counts), ROW_FLOATS, dtype=torch.float32, device=dev)
        dist.all_to_all_single(v_send, v_recv, output_split_sizes=send_counts, input_split_sizes=recv_counts, group=ctx.group)
        grads = None
        xy_grads = []
        off = 0
        for j, (view, rows, offsets, radii, clamped, xy) in enumerate(ctx.per_cam):
            v_rows = v_send[off:off + send_counts[j]]
            off += send_counts[j]
            v_xy = torch.empty(n, 2, dtype=torch.float32, device=dev)
            v_depth = torch.empty(n, dtype=torch.float32, device=dev)
            v_conic = torch.empty(n, 3, dtype=torch.float32, device=dev)
            v_opac = torch.empty(n, dtype=torch.float32, device=dev)
            v_rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
            check(L.b200gs_unpack_rows_grad(n, ptr(radii), ptr(offsets), ptr(v_rows) if v_rows.numel() else None, ptr(v_xy), ptr(v_depth),
                                            ptr(v_conic), None, ptr(v_opac), ptr(v_rgb), st), "b200gs_unpack_rows_grad")
            g = ops.project_backward_raw(view, means, log_scales, raw_quats, ol, shs_dc, shs_rest, ctx.aa, radii, clamped, v_xy, v_depth,
                                         v_conic, v_rgb, v_opac)
            xy_grads.append(v_xy)
            if grads is None:
                grads = list(g)
            else:
                for a, b in zip(grads, g):
                    a.add_(b)
        ctx.xy_grads_out.extend(xy_grads)
        v_means, v_ls, v_q, v_ol, v_dc, v_rest = grads
        return v_means, v_ls, v_q, v_ol.reshape(ctx.opac_shape), v_dc, v_rest, None, None, None, None, None, None


def _sharded_apply(fn, xy_grads_out, *args):
    """apply() with a side list that backward fills with the per-camera mean2D gradients"""
    class _Bound(fn):
        @staticmethod
        def forward(ctx, *a):
            ctx.xy_grads_out = xy_grads_out
            return fn.forward(ctx, *a)
    return _Bound.apply(*args)


class B200DistributedRenderer(torch.nn.Module):
    """Drop-in for ``GSplatDistributedRendererImpl.forward`` (gsplat_distributed_renderer.py:313-414): `pc` holds THIS rank's
    shard; returns this rank's image plus the per-camera projection results the distributed density controller reads
    (``distributed_vanilla_density_controller.py:16-47``)."""

    def __init__(self, anti_aliased: bool = True, group=None, fused: bool = True):
        """fused: when `pc` is the vanilla Gaussian model, run the whole step as one autograd node on the raw parameters
        (_ShardedRasterize: device-side packing, rows consumed in place, one host sync); otherwise the generic path
        below, built from the same ops the single-GPU renderers use."""
        super().__init__()
        self.anti_aliased = anti_aliased
        self.group = group
        self.fused = fused

    def _forward_fused(self, raw, viewpoint_camera, pc, bg_color, scaling_modifier):
        from . import ops
        from ._lib import MODE_GSPLAT
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        dev = bg_color.device
        gathered = torch.empty(world * VIEW_FLOATS, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, pack_view(viewpoint_camera), group=self.group)
        flat = gathered.reshape(world, VIEW_FLOATS).cpu()
        cams = [GatheredView(flat[j], dev) for j in range(world)]
        views = []
        for gv in cams:
            v = ops.make_view(MODE_GSPLAT, gv.width, gv.height, fx=gv.fx, fy=gv.fy, cx=gv.cx, cy=gv.cy, viewmatrix=gv.world_to_camera,
                              campos=gv.camera_center, scale_modifier=scaling_modifier)
            views.append(v)
        xy_grads: List[torch.Tensor] = []
        fn = _ShardedRasterize
        # the per-camera mean2D gradients (what the distributed density controller reads) are appended to this list by backward
        _ShardedRasterize_ctx_hook = xy_grads
        img = _sharded_apply(fn, _ShardedRasterize_ctx_hook, raw["means"], raw["scales"], raw["rotations"], raw["opacities"], raw["shs_dc"],
                             raw["shs_rest"], bg_color, views, rank, self.group, self.anti_aliased, int(pc.active_sh_degree))
        return {
            "render": img.permute(2, 0, 1),
            "cameras": cams,
            "viewspace_points_grads": xy_grads,     # filled by backward: one [n,2] pixel-unit gradient per camera
            "xys_grad_scale_required": True,
        }

    def forward(self, viewpoint_camera, pc, bg_color: torch.Tensor, scaling_modifier=1.0, render_types: list = None, **kwargs):
        from . import ops
        from ._lib import MODE_GSPLAT
        if self.fused:
            from .renderers import _raw_parameters
            raw = _raw_parameters(pc)
            if raw is not None:
                return self._forward_fused(raw, viewpoint_camera, pc, bg_color, scaling_modifier)
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
