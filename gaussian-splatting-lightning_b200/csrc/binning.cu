// K2-K5: tile binning.
//
// The reference backends sort I (tile,Gaussian) pairs on a 64-bit (tile | depth-bits) key: ~6 radix passes over
// 12 B pairs (I*152 B, the largest pure-bandwidth term of the forward).  The same total order — tile
// major, then depth bits, ties in Gaussian-id order (radix sort is stable and pairs are emitted Gaussian-major) — is
// produced here with ~4x less traffic:
//   A. stable radix sort of the N per-Gaussian depth keys (32-bit float bits; culled -> 0xFFFFFFFF)      [N * 16 B * 4 passes]
//   B. scan of tiles-per-Gaussian in depth order -> pair offsets, total I
//   C. emit pairs in depth order: key = tile id (needs only ceil(log2 tiles) bits), value = Gaussian id   [I * 8 B]
//   D. stable radix partition of the pairs by tile id (2 passes at <= 16 bits)                            [I * ~36 B]
//   E. tile ranges from the partitioned keys                                                              [I * 4 B]
// Stable(depth) followed by stable(tile) == stable sort on (tile, depth) — tests compare against the oracle's
// torch.sort(stable) of the 64-bit keys, element for element.
//
// Round 1: the device-wide scan and the two radix sorts are cub:: primitives (header library compiled into this .so);
// the emit / key / range kernels are ours.  Replacing D by a fused emit+partition kernel is the next step.
#include <cub/cub.cuh>

#include "common.cuh"

namespace b200gs {

namespace {

struct LayoutA {
    size_t keys_in, keys_out, ids_in, order, tiles, offsets, temp, temp_bytes, total;
};
struct LayoutB {
    size_t pkeys_in, pkeys_out, pvals_in, temp, temp_bytes, total;
};

struct TilesOfOrder {
    const int32_t* tiles;
    __host__ __device__ int64_t operator()(int32_t g) const { return (int64_t)tiles[g]; }
};

inline int tile_bits_for(int n_tiles) {
    int b = 1;
    while ((1 << b) < n_tiles) ++b;
    return b;
}

struct Taker {
    size_t off = 0;
    size_t operator()(size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

LayoutA make_layout_a(int64_t n) {
    LayoutA L{};
    Taker take;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    L.keys_in = take(nn * 4);
    L.keys_out = take(nn * 4);
    L.ids_in = take(nn * 4);
    L.order = take(nn * 4);
    L.tiles = take(nn * 4);
    L.offsets = take(nn * 8);
    size_t t_sort = 0, t_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)nn, 0, 32);
    cub::TransformInputIterator<int64_t, TilesOfOrder, const int32_t*> it(nullptr, TilesOfOrder{nullptr});
    cub::DeviceScan::InclusiveSum(nullptr, t_scan, it, (int64_t*)nullptr, (int)nn);
    L.temp_bytes = t_sort > t_scan ? t_sort : t_scan;
    L.temp = take(L.temp_bytes);
    L.total = take.off;
    return L;
}

LayoutB make_layout_b(int64_t max_pairs) {
    LayoutB L{};
    Taker take;
    const size_t pp = (size_t)(max_pairs > 0 ? max_pairs : 1);
    L.pkeys_in = take(pp * 4);
    L.pkeys_out = take(pp * 4);
    L.pvals_in = take(pp * 4);
    size_t t_psort = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_psort, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)pp, 0, 16);
    L.temp_bytes = t_psort;
    L.temp = take(t_psort);
    L.total = take.off;
    return L;
}

template <bool GSPLAT>
__global__ void __launch_bounds__(256) depth_keys_kernel(int64_t n, int grid_x, int grid_y, const float2* __restrict__ xy,
                                                         const float* __restrict__ depth, const int32_t* __restrict__ radii,
                                                         uint32_t* __restrict__ keys, int32_t* __restrict__ ids,
                                                         int32_t* __restrict__ tiles) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    uint32_t key = 0xFFFFFFFFu;
    int t = 0;
    if (r > 0) {
        const float2 p = xy[i];
        int x0, y0, x1, y1;
        tile_rect<GSPLAT>(p.x, p.y, (float)r, grid_x, grid_y, x0, y0, x1, y1);
        t = (x1 - x0) * (y1 - y0);
        if (t > 0) key = __float_as_uint(depth[i]);
    }
    keys[i] = key;
    ids[i] = (int32_t)i;
    tiles[i] = t;
}

__global__ void write_total_kernel(int64_t n, const int64_t* __restrict__ offsets, int64_t* __restrict__ d_total) {
    *d_total = n > 0 ? offsets[n - 1] : 0;
}

// One lane per depth-ranked Gaussian; Gaussians with many tiles are written by the whole warp.
template <bool GSPLAT>
__global__ void __launch_bounds__(256) emit_pairs_kernel(int64_t n, int grid_x, int grid_y, int64_t max_pairs,
                                                         const float2* __restrict__ xy, const int32_t* __restrict__ radii,
                                                         const int32_t* __restrict__ order, const int32_t* __restrict__ tiles,
                                                         const int64_t* __restrict__ offsets, uint32_t* __restrict__ pkeys,
                                                         int32_t* __restrict__ pvals) {
    const int64_t rnk = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int g = -1, t = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    int64_t start = 0;
    if (rnk < n) {
        g = order[rnk];
        t = tiles[g];
        if (t > 0) {
            const float2 p = xy[g];
            tile_rect<GSPLAT>(p.x, p.y, (float)radii[g], grid_x, grid_y, x0, y0, x1, y1);
            start = offsets[rnk] - t;
        }
    }
    constexpr int SMALL = 16;
    if (t > 0 && t <= SMALL) {
        const int w = x1 - x0;
        for (int k = 0; k < t; ++k) {
            const int64_t o = start + k;
            if (o < max_pairs) {
                pkeys[o] = (uint32_t)((y0 + k / w) * grid_x + x0 + k % w);
                pvals[o] = g;
            }
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, t > SMALL);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bg = __shfl_sync(0xffffffffu, g, src);
        const int bt = __shfl_sync(0xffffffffu, t, src);
        const int bx0 = __shfl_sync(0xffffffffu, x0, src);
        const int by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bw = __shfl_sync(0xffffffffu, x1, src) - bx0;
        const int64_t bstart = __shfl_sync(0xffffffffu, start, src);
        for (int k = lane; k < bt; k += 32) {
            const int64_t o = bstart + k;
            if (o < max_pairs) {
                pkeys[o] = (uint32_t)((by0 + k / bw) * grid_x + bx0 + k % bw);
                pvals[o] = bg;
            }
        }
    }
}

__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t total, const uint32_t* __restrict__ keys, int2* __restrict__ ranges) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t cur = keys[i];
    if (i == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = keys[i - 1];
        if (prev != cur) {
            ranges[prev].y = (int)i;
            ranges[cur].x = (int)i;
        }
    }
    if (i == total - 1) ranges[cur].y = (int)total;
}

}  // namespace

size_t bin_count_workspace_bytes(int64_t n) { return make_layout_a(n).total; }
size_t bin_sort_workspace_bytes(int64_t, int64_t max_pairs, int, int) { return make_layout_b(max_pairs).total; }

int bin_count(int mode, int width, int height, int64_t n, const float* xy, const float* depth, const int32_t* radii,
              void* ws, size_t ws_bytes, int64_t* d_total, int64_t* host_total, cudaStream_t s) {
    const LayoutA L = make_layout_a(n);
    if (ws_bytes < L.total) {
        set_error("bin_count: workspace too small (%zu < %zu)", ws_bytes, L.total);
        return B200GS_ENOSPACE;
    }
    char* w = (char*)ws;
    uint32_t* keys_in = (uint32_t*)(w + L.keys_in);
    uint32_t* keys_out = (uint32_t*)(w + L.keys_out);
    int32_t* ids_in = (int32_t*)(w + L.ids_in);
    int32_t* order = (int32_t*)(w + L.order);
    int32_t* tiles = (int32_t*)(w + L.tiles);
    int64_t* offsets = (int64_t*)(w + L.offsets);
    const int grid_x = div_up(width, TILE), grid_y = div_up(height, TILE);
    if (n > 0) {
        const unsigned blocks = (unsigned)div_up64(n, 256);
        if (mode == B200GS_MODE_GSPLAT)
            depth_keys_kernel<true><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, (const float2*)xy, depth, radii, keys_in, ids_in, tiles);
        else
            depth_keys_kernel<false><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, (const float2*)xy, depth, radii, keys_in, ids_in, tiles);
        B200GS_LAUNCH_CHECK();
        size_t tb = L.temp_bytes;
        B200GS_CUDA(cub::DeviceRadixSort::SortPairs(w + L.temp, tb, keys_in, keys_out, ids_in, order, (int)n, 0, 32, s));
        cub::TransformInputIterator<int64_t, TilesOfOrder, const int32_t*> it(order, TilesOfOrder{tiles});
        tb = L.temp_bytes;
        B200GS_CUDA(cub::DeviceScan::InclusiveSum(w + L.temp, tb, it, offsets, (int)n, s));
    }
    write_total_kernel<<<1, 1, 0, s>>>(n, offsets, d_total);
    B200GS_LAUNCH_CHECK();
    if (host_total != nullptr) {
        B200GS_CUDA(cudaMemcpyAsync(host_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        B200GS_CUDA(cudaStreamSynchronize(s));
    }
    return B200GS_OK;
}

int bin_sort(int mode, int width, int height, int64_t n, const float* xy, const int32_t* radii, int64_t total,
             int64_t max_pairs, const void* ws_a, void* ws_b, size_t ws_bytes, int32_t* sorted_ids, int32_t* tile_ranges,
             cudaStream_t s) {
    const int grid_x = div_up(width, TILE), grid_y = div_up(height, TILE);
    const int n_tiles = grid_x * grid_y;
    B200GS_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(int32_t) * 2 * (size_t)n_tiles, s));
    if (total > max_pairs) {
        set_error("bin_sort: %lld pairs exceed capacity %lld", (long long)total, (long long)max_pairs);
        return B200GS_ENOSPACE;
    }
    if (total >= (int64_t(1) << 31)) {
        set_error("bin_sort: %lld pairs exceed 2^31", (long long)total);
        return B200GS_ENOSPACE;
    }
    if (total == 0 || n == 0) return B200GS_OK;
    const LayoutA LA = make_layout_a(n);
    const LayoutB L = make_layout_b(max_pairs);
    if (ws_bytes < L.total) {
        set_error("bin_sort: workspace too small (%zu < %zu)", ws_bytes, L.total);
        return B200GS_ENOSPACE;
    }
    const char* wa = (const char*)ws_a;
    char* w = (char*)ws_b;
    const int32_t* order = (const int32_t*)(wa + LA.order);
    const int32_t* tiles = (const int32_t*)(wa + LA.tiles);
    const int64_t* offsets = (const int64_t*)(wa + LA.offsets);
    uint32_t* pkeys_in = (uint32_t*)(w + L.pkeys_in);
    uint32_t* pkeys_out = (uint32_t*)(w + L.pkeys_out);
    int32_t* pvals_in = (int32_t*)(w + L.pvals_in);
    const unsigned blocks = (unsigned)div_up64(n, 256);
    if (mode == B200GS_MODE_GSPLAT)
        emit_pairs_kernel<true><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, max_pairs, (const float2*)xy, radii, order, tiles, offsets, pkeys_in, pvals_in);
    else
        emit_pairs_kernel<false><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, max_pairs, (const float2*)xy, radii, order, tiles, offsets, pkeys_in, pvals_in);
    B200GS_LAUNCH_CHECK();
    size_t tb = L.temp_bytes;
    // temp_b was sized for max_pairs items; cub's requirement is monotone in the item count
    B200GS_CUDA(cub::DeviceRadixSort::SortPairs(w + L.temp, tb, pkeys_in, pkeys_out, pvals_in, sorted_ids, (int)total, 0,
                                                tile_bits_for(n_tiles), s));
    tile_ranges_kernel<<<(unsigned)div_up64(total, 256), 256, 0, s>>>(total, pkeys_out, (int2*)tile_ranges);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

}  // namespace b200gs
