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
    static thread_local int64_t cached_n = -1;
    static thread_local LayoutA cached{};
    if (n == cached_n) return cached;
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
    cached = L;
    cached_n = n;
    return L;
}

LayoutB make_layout_b(int64_t max_pairs) {
    static thread_local int64_t cached_p = -1;
    static thread_local LayoutB cached{};
    if (max_pairs == cached_p) return cached;
    LayoutB L{};
    Taker take;
    const size_t pp = (size_t)(max_pairs > 0 ? max_pairs : 1);
    L.pkeys_in = take(pp * 4);
    L.pkeys_out = take(pp * 4);
    L.pvals_in = take(pp * 4);
    size_t t_psort = 0, t_psort16 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_psort, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)pp, 0, 16);
    cub::DeviceRadixSort::SortPairs(nullptr, t_psort16, (const uint16_t*)nullptr, (uint16_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)pp, 0, 16);
    if (t_psort16 > t_psort) t_psort = t_psort16;
    L.temp_bytes = t_psort;
    L.temp = take(t_psort);
    L.total = take.off;
    cached = L;
    cached_p = max_pairs;
    return L;
}

// ---- exact tile culling ------------------------------------------------------------------------------------------------
// A (tile, splat) pair can only contribute if some pixel sample p of the tile has alpha = o * exp(-q(p - mu)) >= 1/255,
// q(d) = (A dx^2 + C dy^2)/2 + B dx dy, i.e. if the tile's box of pixel samples meets the ellipse E = {q <= ln(255 o)}.
// E and a tile ROW (a band of 16 sample rows) are convex, so the tiles of that row that meet E are exactly those whose
// sample columns meet the x-extent of (E ∩ band): one interval per row, from two clamped evaluations of the ellipse's
// left/right boundary.  Cost O(rows) per Gaussian instead of O(tiles); every pair dropped would have been skipped by
// the blend loop at all 256 pixels, so images and gradients are bit-identical while the pair list (sort, staging,
// blend evaluations) shrinks ~1.8x on the benchmark scene.  The threshold carries a margin for the fp32 / ex2.approx
// rounding of the blend loop, the interval a 0.01 px slack; round-to-nearest intrinsics pin the arithmetic so that the
// counting and the emitting kernel agree on every pair.
// Per-Gaussian inputs of the binning kernels, with element strides: separate contiguous arrays ({2,1,1,3,1}) or columns
// of one [n,12] row buffer ({12,12,12,12,12}).
struct BinSrc {
    const float* xy; const float* depth; const int32_t* radii; const float* conic; const float* opacity;
    int xs, ds, rs, cs, os;
    __device__ __forceinline__ float2 get_xy(int64_t i) const { return *reinterpret_cast<const float2*>(xy + i * xs); }
    __device__ __forceinline__ float get_depth(int64_t i) const { return depth[i * ds]; }
    __device__ __forceinline__ int get_radius(int64_t i) const { return radii[i * rs]; }
};

struct CullE {
    float mx, my, A, B, iA, two_tA, det, ymax, yR;
    int mode;  // 0: full rect (culling off / degenerate conic), 1: spans, 2: nothing visible (opacity <= 1/255)
};

__device__ __forceinline__ CullE load_cull(const float2 p, const BinSrc& src, int64_t g) {
    const float* conic = src.conic ? src.conic + g * src.cs : nullptr;
    const float* opacity = src.opacity ? src.opacity + g * src.os : nullptr;
    CullE e;
    e.mx = p.x; e.my = p.y;
    e.A = 1.f; e.B = 0.f; e.iA = 1.f; e.two_tA = 0.f; e.det = 1.f; e.ymax = 0.f; e.yR = 0.f;
    e.mode = 0;
    if (conic == nullptr) return e;
    const float A = __ldg(conic), B = __ldg(conic + 1), C = __ldg(conic + 2);
    const float o255 = 255.0f * __ldg(opacity);
    if (o255 <= 1.0f) { e.mode = 2; return e; }
    // alpha >= 1/255  <=>  q <= ln(255 o); margin covers the fp32 / ex2.approx rounding of the blend loop
    const float t = __fmaf_rn(__logf(o255), 1.0001f, 1e-3f);
    const float det = __fsub_rn(__fmul_rn(A, C), __fmul_rn(B, B));
    if (!(det > 0.f) || !(A > 0.f) || !(C > 0.f) || !(t < 3.0e38f)) return e;  // degenerate / NaN: keep the full rect
    const float idet = __frcp_rn(det);
    e.A = A; e.B = B; e.iA = __frcp_rn(A);
    e.two_tA = __fmul_rn(__fmul_rn(2.0f, t), A);
    e.det = det;
    e.ymax = __fsqrt_rn(__fmul_rn(e.two_tA, idet));
    const float xext = __fsqrt_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.0f, t), C), idet));
    e.yR = __fmul_rn(__fmul_rn(-B, xext), __frcp_rn(C));  // the ellipse's rightmost point sits at y = yR, leftmost at -yR
    e.mode = 1;
    return e;
}

// tiles [a, b) of tile row ty (clipped to [x0, x1)) whose pixel samples can meet the ellipse
template <bool GSPLAT>
__device__ __forceinline__ bool row_span(const CullE& e, int ty, int x0, int x1, int& a, int& b) {
    a = x0; b = x1;
    if (e.mode == 0) return true;
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float Y0 = __fsub_rn(__fadd_rn(float(ty * TILE), off), e.my);
    const float ya = fmaxf(Y0, -e.ymax), yb = fminf(__fadd_rn(Y0, float(TILE - 1)), e.ymax);
    if (ya > yb) { b = a; return false; }
    const float yr = fminf(yb, fmaxf(ya, e.yR)), yl = fminf(yb, fmaxf(ya, -e.yR));
    const float dr = __fsqrt_rn(fmaxf(0.f, __fmaf_rn(-e.det, __fmul_rn(yr, yr), e.two_tA)));
    const float dl = __fsqrt_rn(fmaxf(0.f, __fmaf_rn(-e.det, __fmul_rn(yl, yl), e.two_tA)));
    const float xr = __fmul_rn(__fadd_rn(__fmul_rn(-e.B, yr), dr), e.iA);   // right end of E ∩ band (relative to mu)
    const float xl = __fmul_rn(__fsub_rn(__fmul_rn(-e.B, yl), dl), e.iA);   // left end
    // tile tx holds sample columns [16 tx + off, 16 tx + off + 15]
    const float inv = 1.0f / float(TILE);
    const float fa = ceilf(__fmul_rn(__fsub_rn(__fadd_rn(xl, e.mx), off + float(TILE - 1) + 0.01f), inv));
    const float fb = floorf(__fmul_rn(__fadd_rn(__fsub_rn(__fadd_rn(xr, e.mx), off), 0.01f), inv));
    if (!(fa <= fb)) {  // also catches NaN
        if (fa == fa && fb == fb) { b = a; return false; }
        return true;    // NaN: keep the whole row
    }
    a = max(x0, (int)fmaxf(fa, -1.0e9f));
    b = min(x1, (int)fminf(fb, 1.0e9f) + 1);
    if (a >= b) { b = a; return false; }
    return true;
}

constexpr int BIG_RECT = 512;   // Gaussians covering more tiles than this are walked by the whole warp

// Shared walk over the rect of one Gaussian per lane.  EMIT=false: returns the number of kept tiles.
// EMIT=true: writes (tile id, g) pairs from `start`, row-major like the reference's emission order.
// Sinks for walk_rect: count only, or stage (tile id, g) pairs of the output window [lo, lo+n) in shared memory.
struct CountSink {
    static constexpr bool kWrites = false;
    __device__ __forceinline__ void put(int64_t, int, int) const {}
};
template <typename KT>
struct StageSink {
    static constexpr bool kWrites = true;
    KT* k; int32_t* v; int64_t lo; int n;
    __device__ __forceinline__ void put(int64_t o, int tile, int g) const {
        const int64_t r = o - lo;
        if (r >= 0 && r < n) { k[r] = (KT)tile; v[r] = g; }
    }
};

// Shared walk over the rect of one Gaussian per lane.  Returns the number of kept tiles; a writing sink receives the
// (tile id, g) pairs at output offsets start, start+1, ... in row-major order (the reference's emission order).
template <bool GSPLAT, typename Sink>
__device__ __forceinline__ int walk_rect(unsigned lane, bool active, int g, const CullE& e, int x0, int y0, int x1, int y1,
                                         int grid_x, int64_t start, const Sink& sink) {
    const int t_rect = active ? (x1 - x0) * (y1 - y0) : 0;
    const bool none = (e.mode == 2);
    int kept = 0;
    if (t_rect > 0 && t_rect <= BIG_RECT && !none) {
        for (int ty = y0; ty < y1; ++ty) {
            int a, b;
            if (!row_span<GSPLAT>(e, ty, x0, x1, a, b)) continue;
            if (Sink::kWrites) {
                for (int tx = a; tx < b; ++tx) sink.put(start + kept + (tx - a), ty * grid_x + tx, g);
            }
            kept += b - a;
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, t_rect > BIG_RECT && !none);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        CullE be;
        be.mx = __shfl_sync(0xffffffffu, e.mx, src); be.my = __shfl_sync(0xffffffffu, e.my, src);
        be.A = __shfl_sync(0xffffffffu, e.A, src); be.B = __shfl_sync(0xffffffffu, e.B, src);
        be.iA = __shfl_sync(0xffffffffu, e.iA, src); be.two_tA = __shfl_sync(0xffffffffu, e.two_tA, src);
        be.det = __shfl_sync(0xffffffffu, e.det, src); be.ymax = __shfl_sync(0xffffffffu, e.ymax, src);
        be.yR = __shfl_sync(0xffffffffu, e.yR, src); be.mode = __shfl_sync(0xffffffffu, e.mode, src);
        const int bg = __shfl_sync(0xffffffffu, g, src);
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bx1 = __shfl_sync(0xffffffffu, x1, src), by1 = __shfl_sync(0xffffffffu, y1, src);
        const int64_t bstart = __shfl_sync(0xffffffffu, start, src);
        int bkept = 0;
        for (int ty = by0; ty < by1; ++ty) {
            int a, b;
            if (!row_span<GSPLAT>(be, ty, bx0, bx1, a, b)) continue;
            if (Sink::kWrites) {
                for (int tx = a + (int)lane; tx < b; tx += 32) sink.put(bstart + bkept + (tx - a), ty * grid_x + tx, bg);
            }
            bkept += b - a;
        }
        if ((int)lane == src) kept = bkept;
    }
    return kept;
}

template <bool GSPLAT>
__global__ void __launch_bounds__(256) depth_keys_kernel(int64_t n, int grid_x, int grid_y, const BinSrc src,
                                                         uint32_t* __restrict__ keys, int32_t* __restrict__ ids,
                                                         int32_t* __restrict__ tiles) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    bool active = false;
    CullE e{};
    if (i < n) {
        const int r = src.get_radius(i);
        if (r > 0) {
            const float2 p = src.get_xy(i);
            tile_rect<GSPLAT>(p.x, p.y, (float)r, grid_x, grid_y, x0, y0, x1, y1);
            e = load_cull(p, src, i);
            active = (x1 - x0) * (y1 - y0) > 0;
        }
    }
    const int t = walk_rect<GSPLAT>(lane, active, (int)i, e, x0, y0, x1, y1, grid_x, 0, CountSink{});
    if (i < n) {
        keys[i] = t > 0 ? __float_as_uint(src.get_depth(i)) : 0xFFFFFFFFu;
        ids[i] = (int32_t)i;
        tiles[i] = t;
    }
}

__global__ void write_total_kernel(int64_t n, const int64_t* __restrict__ offsets, int64_t* __restrict__ d_total) {
    *d_total = n > 0 ? offsets[n - 1] : 0;
}

// One lane per depth-ranked Gaussian.  The 256 consecutive ranks of a block own one contiguous window of the pair arrays
// (offsets are an inclusive scan in depth order), so the pairs are staged in shared memory and written back with
// coalesced stores, EMIT_CHUNK entries at a time (one chunk covers a typical block: 256 x ~13 pairs).
constexpr int EMIT_CHUNK = 4096;

template <bool GSPLAT, typename KT>
__global__ void __launch_bounds__(256) emit_pairs_kernel(int64_t n, int grid_x, int grid_y, int64_t max_pairs, const BinSrc src,
                                                         const int32_t* __restrict__ order, const int32_t* __restrict__ tiles,
                                                         const int64_t* __restrict__ offsets, KT* __restrict__ pkeys,
                                                         int32_t* __restrict__ pvals) {
    __shared__ KT s_k[EMIT_CHUNK];
    __shared__ int32_t s_v[EMIT_CHUNK];
    const int64_t rank0 = int64_t(blockIdx.x) * blockDim.x;
    const int64_t rnk = rank0 + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int g = -1, x0 = 0, y0 = 0, x1 = 0, y1 = 0, t = 0;
    int64_t start = 0;
    CullE e{};
    if (rnk < n) {
        g = order[rnk];
        t = tiles[g];
        if (t > 0) {
            const float2 p = src.get_xy(g);
            tile_rect<GSPLAT>(p.x, p.y, (float)src.get_radius(g), grid_x, grid_y, x0, y0, x1, y1);
            e = load_cull(p, src, g);
            start = offsets[rnk] - t;
        }
    }
    const int64_t last_rank = min(rank0 + (int64_t)blockDim.x, n) - 1;
    const int64_t block_lo = offsets[rank0] - tiles[order[rank0]];
    const int64_t block_hi = min(offsets[last_rank], max_pairs);
    for (int64_t lo = block_lo; lo < block_hi; lo += EMIT_CHUNK) {
        const int cn = (int)min((int64_t)EMIT_CHUNK, block_hi - lo);
        const bool active = (t > 0) && (start < lo + cn) && (start + t > lo);
        walk_rect<GSPLAT>(lane, active, g, e, x0, y0, x1, y1, grid_x, start, StageSink<KT>{s_k, s_v, lo, cn});
        __syncthreads();
        for (int i = threadIdx.x; i < cn; i += blockDim.x) {
            pkeys[lo + i] = s_k[i];
            pvals[lo + i] = s_v[i];
        }
        __syncthreads();
    }
}

// capacity mode: entries [total, cap) of the key buffer get a key above every tile id so they sort to the end
template <typename KT>
__global__ void __launch_bounds__(256) pad_keys_kernel(int64_t cap, const int64_t* __restrict__ d_total, KT* __restrict__ keys, KT pad) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = *d_total + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cap; i += stride) keys[i] = pad;
}

// ranges[t] = [first, last+1) of tile t in the partitioned key array; RV keys per thread (one 8/16-byte load).
template <typename KT>
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t cap, const int64_t* __restrict__ d_total, const KT* __restrict__ keys,
                                                          int2* __restrict__ ranges) {
    constexpr int RV = 16 / sizeof(KT) >= 8 ? 8 : 4;   // 8 x u16 = 16 B, 4 x u32 = 16 B
    const int64_t total = min(cap, *d_total);
    const int64_t i0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * RV;
    if (i0 >= total) return;
    KT k[RV];
    if (i0 + RV <= total) {
        *reinterpret_cast<uint4*>(k) = *reinterpret_cast<const uint4*>(keys + i0);   // keys is 256 B aligned, i0 multiple of RV
    } else {
#pragma unroll
        for (int e = 0; e < RV; ++e) k[e] = (i0 + e < total) ? keys[i0 + e] : KT(0);
    }
    uint32_t prev = (i0 == 0) ? 0xFFFFFFFFu : (uint32_t)keys[i0 - 1];
#pragma unroll
    for (int e = 0; e < RV; ++e) {
        const int64_t i = i0 + e;
        if (i >= total) break;
        const uint32_t cur = k[e];
        if (cur != prev) {
            if (i > 0) ranges[prev].y = (int)i;
            ranges[cur].x = (int)i;
        }
        if (i == total - 1) ranges[cur].y = (int)total;
        prev = cur;
    }
}

template <typename KT>
int emit_sort_ranges(int mode, int64_t n, int grid_x, int grid_y, int n_tiles, const BinSrc& src, bool capacity_mode, int64_t items,
                     const int64_t* d_total, int64_t max_pairs, const int32_t* order, const int32_t* tiles, const int64_t* offsets,
                     void* keys_in, void* keys_out, int32_t* pvals_in, void* temp, size_t temp_bytes, int32_t* sorted_ids,
                     int32_t* tile_ranges, cudaStream_t s) {
    KT* kin = (KT*)keys_in;
    KT* kout = (KT*)keys_out;
    const unsigned blocks = (unsigned)div_up64(n, 256);
    if (mode == B200GS_MODE_GSPLAT)
        emit_pairs_kernel<true, KT><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, max_pairs, src, order, tiles, offsets, kin, pvals_in);
    else
        emit_pairs_kernel<false, KT><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, max_pairs, src, order, tiles, offsets, kin, pvals_in);
    B200GS_LAUNCH_CHECK();
    int bits = tile_bits_for(n_tiles);
    if (capacity_mode) {
        pad_keys_kernel<KT><<<512, 256, 0, s>>>(items, d_total, kin, (KT)(1u << bits));
        B200GS_LAUNCH_CHECK();
        bits += 1;
    }
    size_t tb = temp_bytes;
    // temp was sized for max_pairs 32-bit keys; cub's requirement is monotone in the item count and key width
    B200GS_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, kin, kout, pvals_in, sorted_ids, (int)items, 0, bits, s));
    constexpr int RV = 16 / sizeof(KT) >= 8 ? 8 : 4;
    tile_ranges_kernel<KT><<<(unsigned)div_up64(div_up64(items, RV), 256), 256, 0, s>>>(items, d_total, kout, (int2*)tile_ranges);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

}  // namespace

size_t bin_count_workspace_bytes(int64_t n) { return make_layout_a(n).total; }
size_t bin_sort_workspace_bytes(int64_t, int64_t max_pairs, int, int) { return make_layout_b(max_pairs).total; }

static BinSrc make_src(int row_stride, const float* xy, const float* depth, const int32_t* radii, const float* conic, const float* opacity) {
    if (row_stride > 0) return BinSrc{xy, depth, radii, conic, opacity, row_stride, row_stride, row_stride, row_stride, row_stride};
    return BinSrc{xy, depth, radii, conic, opacity, 2, 1, 1, 3, 1};
}

int bin_count(int mode, int width, int height, int64_t n, int row_stride, const float* xy, const float* depth, const int32_t* radii,
              const float* conic, const float* opacity, void* ws, size_t ws_bytes, int64_t* d_total, int64_t* host_total,
              int sync_host, cudaStream_t s) {
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
        const BinSrc src = make_src(row_stride, xy, depth, radii, conic, opacity);
        if (mode == B200GS_MODE_GSPLAT)
            depth_keys_kernel<true><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, src, keys_in, ids_in, tiles);
        else
            depth_keys_kernel<false><<<blocks, 256, 0, s>>>(n, grid_x, grid_y, src, keys_in, ids_in, tiles);
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
        if (sync_host) B200GS_CUDA(cudaStreamSynchronize(s));
    }
    return B200GS_OK;
}

int bin_sort(int mode, int width, int height, int64_t n, int row_stride, const float* xy, const int32_t* radii, const float* conic,
             const float* opacity, int64_t total, const int64_t* d_total, int64_t max_pairs, const void* ws_a, void* ws_b,
             size_t ws_bytes, int32_t* sorted_ids, int32_t* tile_ranges, cudaStream_t s) {
    const int grid_x = div_up(width, TILE), grid_y = div_up(height, TILE);
    const int n_tiles = grid_x * grid_y;
    const bool capacity_mode = total < 0;   // pair count known on the device only: sort the whole capacity, padded
    B200GS_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(int32_t) * 2 * (size_t)n_tiles, s));
    if (!capacity_mode && total > max_pairs) {
        set_error("bin_sort: %lld pairs exceed capacity %lld", (long long)total, (long long)max_pairs);
        return B200GS_ENOSPACE;
    }
    const int64_t items = capacity_mode ? max_pairs : total;
    if (items >= (int64_t(1) << 31)) {
        set_error("bin_sort: %lld pairs exceed 2^31", (long long)items);
        return B200GS_ENOSPACE;
    }
    if (items == 0 || n == 0) return B200GS_OK;
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
    const BinSrc src = make_src(row_stride, xy, nullptr, radii, conic, opacity);
    const int bits = tile_bits_for(n_tiles) + (capacity_mode ? 1 : 0);
    if (bits <= 16)   // every image up to ~4K: 16-bit tile keys -> 12 B instead of 16 B per pair and radix pass
        return emit_sort_ranges<uint16_t>(mode, n, grid_x, grid_y, n_tiles, src, capacity_mode, items, d_total, max_pairs, order, tiles,
                                          offsets, pkeys_in, pkeys_out, pvals_in, w + L.temp, L.temp_bytes, sorted_ids, tile_ranges, s);
    return emit_sort_ranges<uint32_t>(mode, n, grid_x, grid_y, n_tiles, src, capacity_mode, items, d_total, max_pairs, order, tiles,
                                      offsets, pkeys_in, pkeys_out, pvals_in, w + L.temp, L.temp_bytes, sorted_ids, tile_ranges, s);
}

// ---- row packing for the Gaussian-sharded exchange -------------------------------------------------------------------
namespace {

struct VisibleFlag {
    const int32_t* radii;
    __host__ __device__ int32_t operator()(int32_t i) const { return radii[i] > 0 ? 1 : 0; }
};

__global__ void __launch_bounds__(256) pack_rows_kernel(int64_t n, const float2* __restrict__ xy, const float* __restrict__ depth,
                                                        const float* __restrict__ conic, const float* __restrict__ comp,
                                                        const float* __restrict__ opacity, const float* __restrict__ rgb,
                                                        const int32_t* __restrict__ radii, const int32_t* __restrict__ offsets,
                                                        float* __restrict__ rows, int64_t* __restrict__ d_count) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = radii[i];
    const int o = offsets[i];
    if (i == n - 1) *d_count = o + (r > 0 ? 1 : 0);
    if (r <= 0) return;
    float4* out = reinterpret_cast<float4*>(rows + int64_t(o) * B200GS_ROW_FLOATS);
    const float2 p = xy[i];
    out[0] = make_float4(p.x, p.y, depth[i], conic[3 * i]);
    out[1] = make_float4(conic[3 * i + 1], conic[3 * i + 2], comp ? comp[i] : 1.0f, opacity[i]);
    out[2] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], __int_as_float(r));
}

__global__ void __launch_bounds__(256) unpack_rows_grad_kernel(int64_t n, const int32_t* __restrict__ radii,
                                                               const int32_t* __restrict__ offsets, const float* __restrict__ v_rows,
                                                               float2* __restrict__ v_xy, float* __restrict__ v_depth,
                                                               float* __restrict__ v_conic, float* __restrict__ v_comp,
                                                               float* __restrict__ v_opacity, float* __restrict__ v_rgb) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
    if (radii[i] > 0) {
        const float4* in = reinterpret_cast<const float4*>(v_rows + int64_t(offsets[i]) * B200GS_ROW_FLOATS);
        a = in[0]; b = in[1]; c = in[2];
    }
    v_xy[i] = make_float2(a.x, a.y);
    v_depth[i] = a.z;
    v_conic[3 * i] = a.w; v_conic[3 * i + 1] = b.x; v_conic[3 * i + 2] = b.y;
    if (v_comp) v_comp[i] = b.z;
    v_opacity[i] = b.w;
    v_rgb[3 * i] = c.x; v_rgb[3 * i + 1] = c.y; v_rgb[3 * i + 2] = c.z;
}

}  // namespace

size_t pack_rows_workspace_bytes(int64_t n) {
    size_t t = 0;
    cub::TransformInputIterator<int32_t, VisibleFlag, cub::CountingInputIterator<int32_t>> it(cub::CountingInputIterator<int32_t>(0),
                                                                                                VisibleFlag{nullptr});
    cub::DeviceScan::ExclusiveSum(nullptr, t, it, (int32_t*)nullptr, (int)(n > 0 ? n : 1));
    return align_up(t, 256);
}

int pack_rows(int64_t n, const float* xy, const float* depth, const float* conic, const float* comp, const float* opacity,
              const float* rgb, const int32_t* radii, void* ws, size_t ws_bytes, int32_t* offsets, float* rows, int64_t* d_count,
              cudaStream_t s) {
    if (n == 0) {
        B200GS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), s));
        return B200GS_OK;
    }
    cub::TransformInputIterator<int32_t, VisibleFlag, cub::CountingInputIterator<int32_t>> it(cub::CountingInputIterator<int32_t>(0),
                                                                                                VisibleFlag{radii});
    size_t tb = ws_bytes;
    B200GS_CUDA(cub::DeviceScan::ExclusiveSum(ws, tb, it, offsets, (int)n, s));
    pack_rows_kernel<<<(unsigned)div_up64(n, 256), 256, 0, s>>>(n, (const float2*)xy, depth, conic, comp, opacity, rgb, radii, offsets,
                                                               rows, d_count);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

int unpack_rows_grad(int64_t n, const int32_t* radii, const int32_t* offsets, const float* v_rows, float* v_xy, float* v_depth,
                     float* v_conic, float* v_comp, float* v_opacity, float* v_rgb, cudaStream_t s) {
    if (n == 0) return B200GS_OK;
    unpack_rows_grad_kernel<<<(unsigned)div_up64(n, 256), 256, 0, s>>>(n, radii, offsets, v_rows, (float2*)v_xy, v_depth, v_conic, v_comp,
                                                                      v_opacity, v_rgb);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

}  // namespace b200gs
