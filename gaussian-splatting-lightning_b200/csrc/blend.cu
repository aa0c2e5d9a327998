// K6 / K7: per-tile front-to-back alpha compositing, forward and backward.
//
// Restates the published per-pixel loops of diff-gaussian-rasterization@59f5f77 renderCUDA (vanilla mode) and gsplat's
// rasterize_to_pixels (gsplat mode) — SURVEY.md §8c / Appendix B; the sources are not under /root/reference, the CPU
// restatement in oracle/gs_oracle.py::blend is the checker.
//
// One CTA per 16x16 tile, one thread per pixel; a warp covers an 8x4 pixel block so that a splat's footprint skips
// whole warps (warp-ballot early out).  Each round stages up to 256 splats of the tile's depth-sorted slab into shared
// memory (coalesced id read, L2-resident gathers of the 36 B splat record), then every pixel walks the staged slab.
// Backward walks the slab in reverse, only up to the deepest contributor of the tile, reduces each splat's 9 partials
// across the warp with shuffles and issues one atomic per (warp, splat, component) — 32x fewer L2 atomics than the
// reference's per-pixel atomicAdd.
#include "common.cuh"

namespace b200gs {

namespace {

constexpr int BLOCK_PIX = TILE * TILE;  // 256 threads
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_STOP = 1e-4f;

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
}

template <int CH, bool GSPLAT>
__global__ void __launch_bounds__(BLOCK_PIX) blend_fwd_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const float2* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              float* __restrict__ image, int64_t pix_stride, int64_t ch_stride,
                                                              float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                                                              float* __restrict__ alpha_out) {
    __shared__ float2 s_xy[BLOCK_PIX];
    __shared__ float4 s_co[BLOCK_PIX];
    __shared__ float s_col[CH][BLOCK_PIX];

    const int tid = threadIdx.x;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float pxf = float(px) + (GSPLAT ? 0.5f : 0.0f);
    const float pyf = float(py) + (GSPLAT ? 0.5f : 0.0f);
    const float amax = GSPLAT ? 0.999f : 0.99f;

    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    bool done = !inside;
    float T = 1.0f;
    int last = 0;
    float C[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) C[c] = 0.f;

    for (int base = 0; todo > 0; base += BLOCK_PIX, todo -= BLOCK_PIX) {
        if (__syncthreads_and(done)) break;
        const int cnt = min(BLOCK_PIX, todo);
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + base + tid);
            s_xy[tid] = __ldg(xy + g);
            s_co[tid] = make_float4(__ldg(conic + 3 * g), __ldg(conic + 3 * g + 1), __ldg(conic + 3 * g + 2), __ldg(opacity + g));
#pragma unroll
            for (int c = 0; c < CH; ++c) s_col[c][tid] = __ldg(colors + int64_t(g) * CH + c);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            if (__all_sync(0xffffffffu, done)) break;
            if (done) continue;
            const float2 m = s_xy[j];
            const float4 co = s_co[j];
            const float dx = m.x - pxf, dy = m.y - pyf;
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float a = fminf(amax, co.w * __expf(power));
            if (a < ALPHA_MIN) continue;
            const float nT = T * (1.0f - a);
            if (GSPLAT ? (nT <= T_STOP) : (nT < T_STOP)) {
                done = true;
                continue;
            }
            const float w = a * T;
#pragma unroll
            for (int c = 0; c < CH; ++c) C[c] += s_col[c][j] * w;
            T = nT;
            last = base + j + 1;
        }
    }
    if (inside) {
        const int64_t pix = int64_t(py) * width + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        if (alpha_out) alpha_out[pix] = 1.0f - T;
#pragma unroll
        for (int c = 0; c < CH; ++c) image[pix * pix_stride + c * ch_stride] = C[c] + (bg ? T * __ldg(bg + c) : 0.f);
    }
}

template <int CH, bool GSPLAT>
__global__ void __launch_bounds__(BLOCK_PIX) blend_bwd_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const float2* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
                                                              const float* __restrict__ v_image, int64_t pix_stride, int64_t ch_stride,
                                                              const float* __restrict__ v_alpha, float sx, float sy,
                                                              float* __restrict__ v_xy, float* __restrict__ v_conic,
                                                              float* __restrict__ v_opacity, float* __restrict__ v_colors,
                                                              float* __restrict__ v_xy_abs) {
    __shared__ float2 s_xy[BLOCK_PIX];
    __shared__ float4 s_co[BLOCK_PIX];
    __shared__ float s_col[CH][BLOCK_PIX];
    __shared__ int s_id[BLOCK_PIX];
    __shared__ int s_wmax[BLOCK_PIX / 32];

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float pxf = float(px) + (GSPLAT ? 0.5f : 0.0f);
    const float pyf = float(py) + (GSPLAT ? 0.5f : 0.0f);
    const float amax = GSPLAT ? 0.999f : 0.99f;
    const int64_t pix = int64_t(py) * width + px;

    const int2 range = ranges[tile];
    const float Tf = inside ? final_T[pix] : 0.f;
    const int last = inside ? n_contrib[pix] : 0;
    float vo[CH];
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        vo[c] = inside ? __ldg(v_image + pix * pix_stride + c * ch_stride) : 0.f;
        if (bg) bg_dot += __ldg(bg + c) * vo[c];
    }
    const float va = (v_alpha && inside) ? __ldg(v_alpha + pix) : 0.f;

    const int wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) s_wmax[tid >> 5] = wmax;
    __syncthreads();
    int max_last = 0;
#pragma unroll
    for (int w = 0; w < BLOCK_PIX / 32; ++w) max_last = max(max_last, s_wmax[w]);
    if (max_last == 0) return;

    float T = Tf;
    float buf[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) buf[c] = 0.f;

    for (int hi = max_last; hi > 0; hi -= BLOCK_PIX) {
        const int lo = max(0, hi - BLOCK_PIX);
        const int cnt = hi - lo;
        __syncthreads();
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + lo + tid);
            s_id[tid] = g;
            s_xy[tid] = __ldg(xy + g);
            s_co[tid] = make_float4(__ldg(conic + 3 * g), __ldg(conic + 3 * g + 1), __ldg(conic + 3 * g + 2), __ldg(opacity + g));
#pragma unroll
            for (int c = 0; c < CH; ++c) s_col[c][tid] = __ldg(colors + int64_t(g) * CH + c);
        }
        __syncthreads();
        if (wmax <= lo) continue;  // this warp has no contributor in the batch
        for (int j = min(cnt, wmax - lo) - 1; j >= 0; --j) {
            bool valid = (lo + j) < last;
            float dx = 0.f, dy = 0.f, G = 0.f, a = 0.f;
            float4 co = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float2 m = s_xy[j];
                co = s_co[j];
                dx = m.x - pxf; dy = m.y - pyf;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                G = __expf(power);
                a = fminf(amax, co.w * G);
                valid = !(power > 0.0f) && (a >= ALPHA_MIN);
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float g_x = 0.f, g_y = 0.f, g_a = 0.f, g_b = 0.f, g_c = 0.f, g_o = 0.f;
            float g_col[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) g_col[c] = 0.f;
            if (valid) {
                const float ra = 1.0f / (1.0f - a);
                T *= ra;
                const float fac = a * T;
                float v_al = 0.f;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float col = s_col[c][j];
                    g_col[c] = fac * vo[c];
                    v_al += (col * T - buf[c] * ra) * vo[c];
                    buf[c] += col * fac;
                }
                v_al += Tf * ra * (va - bg_dot);
                if (!GSPLAT || (co.w * G <= 0.999f)) {
                    const float v_sigma = -co.w * G * v_al;
                    g_a = 0.5f * v_sigma * dx * dx;
                    g_b = v_sigma * dx * dy;
                    g_c = 0.5f * v_sigma * dy * dy;
                    g_x = v_sigma * (co.x * dx + co.y * dy);
                    g_y = v_sigma * (co.y * dx + co.z * dy);
                    g_o = G * v_al;
                }
            }
            float ax = 0.f, ay = 0.f;
            if (v_xy_abs) {
                ax = warp_sum(fabsf(g_x));
                ay = warp_sum(fabsf(g_y));
            }
            g_x = warp_sum(g_x); g_y = warp_sum(g_y);
            g_a = warp_sum(g_a); g_b = warp_sum(g_b); g_c = warp_sum(g_c);
            g_o = warp_sum(g_o);
#pragma unroll
            for (int c = 0; c < CH; ++c) g_col[c] = warp_sum(g_col[c]);
            if (lane == 0) {
                const int g = s_id[j];
                atomicAdd(v_xy + 2 * g, g_x * sx);
                atomicAdd(v_xy + 2 * g + 1, g_y * sy);
                atomicAdd(v_conic + 3 * g, g_a);
                atomicAdd(v_conic + 3 * g + 1, g_b);
                atomicAdd(v_conic + 3 * g + 2, g_c);
                atomicAdd(v_opacity + g, g_o);
#pragma unroll
                for (int c = 0; c < CH; ++c) atomicAdd(v_colors + int64_t(g) * CH + c, g_col[c]);
                if (v_xy_abs) {
                    atomicAdd(v_xy_abs + 2 * g, ax);
                    atomicAdd(v_xy_abs + 2 * g + 1, ay);
                }
            }
        }
    }
}

template <int CH>
int fwd_dispatch(int mode, int width, int height, const int32_t* ranges, const int32_t* ids, const float* xy, const float* conic,
                 const float* opacity, const float* colors, const float* bg, float* image, int64_t ps, int64_t cs, float* final_T,
                 int32_t* n_contrib, float* alpha, cudaStream_t s) {
    const int gx = div_up(width, TILE), gy = div_up(height, TILE);
    dim3 grid(gx, gy);
    if (mode == B200GS_MODE_GSPLAT)
        blend_fwd_kernel<CH, true><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, (const float2*)xy, conic,
                                                              opacity, colors, bg, image, ps, cs, final_T, n_contrib, alpha);
    else
        blend_fwd_kernel<CH, false><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, (const float2*)xy, conic,
                                                               opacity, colors, bg, image, ps, cs, final_T, n_contrib, alpha);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

template <int CH>
int bwd_dispatch(int mode, int width, int height, const int32_t* ranges, const int32_t* ids, const float* xy, const float* conic,
                 const float* opacity, const float* colors, const float* bg, const float* final_T, const int32_t* n_contrib,
                 const float* v_image, int64_t ps, int64_t cs, const float* v_alpha, float sx, float sy, float* v_xy, float* v_conic,
                 float* v_opacity, float* v_colors, float* v_xy_abs, cudaStream_t s) {
    const int gx = div_up(width, TILE), gy = div_up(height, TILE);
    dim3 grid(gx, gy);
    if (mode == B200GS_MODE_GSPLAT)
        blend_bwd_kernel<CH, true><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, (const float2*)xy, conic,
                                                              opacity, colors, bg, final_T, n_contrib, v_image, ps, cs, v_alpha, sx,
                                                              sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs);
    else
        blend_bwd_kernel<CH, false><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, (const float2*)xy, conic,
                                                               opacity, colors, bg, final_T, n_contrib, v_image, ps, cs, v_alpha, sx,
                                                               sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

}  // namespace

int launch_blend_fwd(int mode, int width, int height, int channels, const int32_t* ranges, const int32_t* ids, const float* xy,
                     const float* conic, const float* opacity, const float* colors, const float* bg, float* image,
                     int64_t pix_stride, int64_t ch_stride, float* final_T, int32_t* n_contrib, float* alpha, cudaStream_t s) {
    switch (channels) {
        case 1: return fwd_dispatch<1>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, s);
        case 2: return fwd_dispatch<2>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, s);
        case 3: return fwd_dispatch<3>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, s);
        case 4: return fwd_dispatch<4>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, s);
    }
    set_error("blend_fwd: unsupported channel count %d (1..4)", channels);
    return B200GS_EINVAL;
}

int launch_blend_bwd(int mode, int width, int height, int channels, const int32_t* ranges, const int32_t* ids, const float* xy,
                     const float* conic, const float* opacity, const float* colors, const float* bg, const float* final_T,
                     const int32_t* n_contrib, const float* v_image, int64_t pix_stride, int64_t ch_stride, const float* v_alpha,
                     float sx, float sy, float* v_xy, float* v_conic, float* v_opacity, float* v_colors, float* v_xy_abs,
                     cudaStream_t s) {
    switch (channels) {
        case 1: return bwd_dispatch<1>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 2: return bwd_dispatch<2>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 3: return bwd_dispatch<3>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 4: return bwd_dispatch<4>(mode, width, height, ranges, ids, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
    }
    set_error("blend_bwd: unsupported channel count %d (1..4)", channels);
    return B200GS_EINVAL;
}

}  // namespace b200gs
