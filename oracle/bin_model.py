"""CPU model of the hierarchical binning of libb200gs (csrc/binning.cu) — TEST INFRASTRUCTURE ONLY.

Nothing under gaussian-splatting-lightning_b200/ may import this module; it exists so that the *algorithm* of K2-K5 — stable
depth sort, stable partition of (8x8-tile cell, Gaussian) pairs carrying 64-bit tile masks, order-preserving multi-split of
every cell's list in 256-entry chunks (per-chunk tile counts, per-cell chunk prefixes, one scan over the tiles, scatter at
tile start + chunk prefix + rank inside the chunk) — is checked on CPU against the reference order, i.e. the stable sort of the
(tile << 32 | depth bits) keys that the reference backends perform (gs_oracle.build_sort_keys / sort_and_ranges, which
restate internal/utils/gaussian_projection.py:173-208), without a GPU.  Plain loops: small scenes only.
"""
from typing import Optional, Tuple

import torch

SUPER = 8          # coarse cell edge in tiles (binning.cu: SUPER)
CHUNK = 256        # list entries per block of the count / scatter kernels (binning.cu: CHUNK)


def hierarchical_binning(depth: torch.Tensor, rect_min: torch.Tensor, rect_max: torch.Tensor, grid_x: int, grid_y: int,
                         keep: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """depth [N] fp32, rect_min/rect_max [N,2] tile rects ([min, max) per axis; empty rect = culled Gaussian).
    keep: optional bool [N, grid_y, grid_x] — which tiles of its rect a Gaussian keeps (exact tile culling); None = all.
    Returns (sorted_ids int32 [I], ranges int64 [n_tiles, 2]) with empty tiles reported as (0, 0) like binning.cu."""
    n = depth.shape[0]
    cgx, cgy = (grid_x + SUPER - 1) // SUPER, (grid_y + SUPER - 1) // SUPER
    x0, y0, x1, y1 = rect_min[:, 0].tolist(), rect_min[:, 1].tolist(), rect_max[:, 0].tolist(), rect_max[:, 1].tolist()
    area = [(max(0, x1[i] - x0[i])) * (max(0, y1[i] - y0[i])) for i in range(n)]

    # A: depth keys (float bits; off-screen last), stable sort
    bits = depth.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = torch.where(torch.tensor(area) > 0, bits, torch.full_like(bits, 0xFFFFFFFF))
    order = torch.sort(keys, stable=True).indices.tolist()

    # B: (cell, {mask, id}) pairs in depth order, cells row-major; stable partition by cell
    cell_lists = [[] for _ in range(cgx * cgy)]
    for g in order:
        if area[g] == 0:
            continue
        for cy in range(y0[g] // SUPER, (y1[g] - 1) // SUPER + 1):
            for cx in range(x0[g] // SUPER, (x1[g] - 1) // SUPER + 1):
                mask = 0
                for ty in range(max(y0[g], cy * SUPER), min(y1[g], cy * SUPER + SUPER)):
                    for tx in range(max(x0[g], cx * SUPER), min(x1[g], cx * SUPER + SUPER)):
                        if keep is None or bool(keep[g, ty, tx]):
                            mask |= 1 << ((ty - cy * SUPER) * SUPER + (tx - cx * SUPER))
                cell_lists[cy * cgx + cx].append((mask, g))      # pairs with an empty mask stay in the list, like on the GPU

    # C: per-chunk tile counts;  D: chunk prefixes per (cell, tile), tile totals, scan in tile-id order
    n_tiles = grid_x * grid_y
    tile_total = [0] * n_tiles
    chunk_pre = {}
    for cell, lst in enumerate(cell_lists):
        run = [0] * (SUPER * SUPER)
        for c0 in range(0, len(lst), CHUNK):
            chunk_pre[(cell, c0)] = list(run)
            for mask, _ in lst[c0:c0 + CHUNK]:
                for t in range(SUPER * SUPER):
                    run[t] += (mask >> t) & 1
        cx, cy = cell % cgx, cell // cgx
        for t in range(SUPER * SUPER):
            tx, ty = cx * SUPER + (t % SUPER), cy * SUPER + (t // SUPER)
            if tx < grid_x and ty < grid_y:
                tile_total[ty * grid_x + tx] = run[t]
            else:
                assert run[t] == 0
    tile_start, acc = [], 0
    for v in tile_total:
        tile_start.append(acc)
        acc += v

    # E: scatter — position = tile start + chunk prefix + rank of the entry among the chunk's entries with that tile bit
    out = [-1] * acc
    for cell, lst in enumerate(cell_lists):
        cx, cy = cell % cgx, cell // cgx
        for c0 in range(0, len(lst), CHUNK):
            rank = [0] * (SUPER * SUPER)
            pre = chunk_pre[(cell, c0)]
            for mask, g in lst[c0:c0 + CHUNK]:
                for t in range(SUPER * SUPER):
                    if (mask >> t) & 1:
                        tile = (cy * SUPER + t // SUPER) * grid_x + cx * SUPER + t % SUPER
                        pos = tile_start[tile] + pre[t] + rank[t]
                        assert out[pos] == -1
                        out[pos] = g
                        rank[t] += 1
    assert all(v >= 0 for v in out)
    ranges = torch.tensor([[s, s + c] if c > 0 else [0, 0] for s, c in zip(tile_start, tile_total)], dtype=torch.int64).reshape(n_tiles, 2)
    return torch.tensor(out, dtype=torch.int32), ranges
