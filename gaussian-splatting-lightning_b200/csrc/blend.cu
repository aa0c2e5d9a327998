// K6 / K7: per-tile front-to-back alpha compositing, forward and backward.
//
// Restates the published per-pixel loops of diff-gaussian-rasterization@59f5f77 renderCUDA (vanilla mode) and gsplat's
// rasterize_to_pixels (gsplat mode) — SURVEY.md §8c / Appendix B; the sources are not under /root/reference, the CPU
// restatement in oracle/gs_oracle.py::blend is the checker.
//
// One CTA per 16x16 tile, one thread per pixel; a warp covers an 8x4 pixel block so that a splat's footprint skips
// whole warps (warp-ballot early out).  Each round stages up to 256 splats of the tile's depth-sorted slab into shared
// memory (coalesced id read, L2-resident gathers of the 36 B splat record), then every pixel walks the staged slab.
//
// Forward: the conic is pre-scaled by -0.5*log2(e) / -log2(e) while staging, so the per-(pixel,splat) body is
// 5 FP32 ops + one MUFU.EX2 + compare/blend; colours are fetched (one LDS.128) only by contributing lanes.
//
// Backward walks the slab in reverse, only up to the deepest contributor of the tile.  Each pixel evaluates RB
// consecutive splats and keeps their 9 partial derivatives in registers; the warp then reduces them with a
// reduce-SCATTER butterfly (halving exchanges: RB -> RB/2 -> ... -> 1 value per lane), i.e. ~10 shuffles per splat
// instead of 45 for nine independent all-reduces, and 32x fewer L2 atomics than the reference's per-pixel atomicAdd.
#include "common.cuh"

namespace b200gs {

namespace {

constexpr int BLOCK_PIX = TILE * TILE;  // 256 threads
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_STOP = 1e-4f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int CH, bool GSPLAT>
__global__ void __launch_bounds__(BLOCK_PIX) blend_fwd_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const float2* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              float* __restrict__ image, int64_t pix_stride, int64_t ch_stride,
                                                              float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                                                              float* __restrict__ alpha_out) {
    __shared__ float4 s_g1[BLOCK_PIX];  // x, y, -0.5*log2e*A, -log2e*B
    __shared__ float2 s_g2[BLOCK_PIX];  // -0.5*log2e*C, opacity
    __shared__ float4 s_col[BLOCK_PIX];

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
            const float2 m = __ldg(xy + g);
            s_g1[tid] = make_float4(m.x, m.y, (-0.5f * LOG2E) * __ldg(conic + 3 * g), -LOG2E * __ldg(conic + 3 * g + 1));
            s_g2[tid] = make_float2((-0.5f * LOG2E) * __ldg(conic + 3 * g + 2), __ldg(opacity + g));
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * CH + c);
            s_col[tid] = make_float4(col[0], col[1], col[2], col[3]);
        }
        __syncthreads();
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            if (__all_sync(FULL, done)) break;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < cnt && !done) {
                    const float4 g1 = s_g1[j];
                    const float2 g2 = s_g2[j];
                    const float dx = g1.x - pxf, dy = g1.y - pyf;
                    // power * log2(e) = a' dx^2 + b' dx dy + c' dy^2
                    const float p2 = fmaf(g2.x * dy, dy, fmaf(g1.w, dy, g1.z * dx) * dx);
                    const float a = fminf(amax, g2.y * ex2_approx(p2));
                    if (!(p2 > 0.0f) && !(a < ALPHA_MIN)) {
                        const float nT = fmaf(-a, T, T);
                        if (GSPLAT ? (nT <= T_STOP) : (nT < T_STOP)) {
                            done = true;
                        } else {
                            const float w = a * T;
                            const float4 col = s_col[j];
                            C[0] = fmaf(col.x, w, C[0]);
                            if (CH > 1) C[1 % CH] = fmaf(col.y, w, C[1 % CH]);
                            if (CH > 2) C[2 % CH] = fmaf(col.z, w, C[2 % CH]);
                            if (CH > 3) C[3 % CH] = fmaf(col.w, w, C[3 % CH]);
                            T = nT;
                            last = base + j + 1;
                        }
                    }
                }
            }
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

// ---- warp reduce-scatter: RB values per lane -> lane keeps the 32-lane total of slot rs_slot(lane) ---------------------
template <int RB>
__device__ __forceinline__ int rs_slot(unsigned lane);
template <>
__device__ __forceinline__ int rs_slot<8>(unsigned lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
template <>
__device__ __forceinline__ int rs_slot<4>(unsigned lane) { return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); }

__device__ __forceinline__ float rs_step(float lo, float hi, bool up, int mask) {
    const float send = up ? lo : hi;
    const float keep = up ? hi : lo;
    return keep + __shfl_xor_sync(FULL, send, mask);
}

template <int RB>
__device__ __forceinline__ float reduce_scatter(const float* p, unsigned lane);

template <>
__device__ __forceinline__ float reduce_scatter<8>(const float* p, unsigned lane) {
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
    const float q0 = rs_step(p[0], p[4], b16, 16), q1 = rs_step(p[1], p[5], b16, 16);
    const float q2 = rs_step(p[2], p[6], b16, 16), q3 = rs_step(p[3], p[7], b16, 16);
    const float r0 = rs_step(q0, q2, b8, 8), r1 = rs_step(q1, q3, b8, 8);
    float s = rs_step(r0, r1, b4, 4);
    s += __shfl_xor_sync(FULL, s, 2);
    s += __shfl_xor_sync(FULL, s, 1);
    return s;
}

template <>
__device__ __forceinline__ float reduce_scatter<4>(const float* p, unsigned lane) {
    const bool b16 = lane & 16, b8 = lane & 8;
    const float q0 = rs_step(p[0], p[2], b16, 16), q1 = rs_step(p[1], p[3], b16, 16);
    float s = rs_step(q0, q1, b8, 8);
    s += __shfl_xor_sync(FULL, s, 4);
    s += __shfl_xor_sync(FULL, s, 2);
    s += __shfl_xor_sync(FULL, s, 1);
    return s;
}

template <int CH, bool GSPLAT, bool ABS, int RB>
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
    constexpr int NT = 6 + CH + (ABS ? 2 : 0);  // x y a b c o colours [|x| |y|]
    __shared__ float4 s_g1[BLOCK_PIX];  // x, y, A, B
    __shared__ float2 s_g2[BLOCK_PIX];  // C, opacity
    __shared__ float4 s_col[BLOCK_PIX];
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
    const float tail = Tf * (va - bg_dot);  // d(out)/d(alpha_i) term through everything behind the last contributor

    const int wmax = __reduce_max_sync(FULL, last);
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
    const int my_slot = rs_slot<RB>(lane);
    const bool writer = (lane & (32 / RB - 1)) == 0;

    for (int hi = max_last; hi > 0; hi -= BLOCK_PIX) {
        const int lo = max(0, hi - BLOCK_PIX);
        const int cnt = hi - lo;
        __syncthreads();
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + lo + tid);
            s_id[tid] = g;
            const float2 m = __ldg(xy + g);
            s_g1[tid] = make_float4(m.x, m.y, __ldg(conic + 3 * g), __ldg(conic + 3 * g + 1));
            s_g2[tid] = make_float2(__ldg(conic + 3 * g + 2), __ldg(opacity + g));
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * CH + c);
            s_col[tid] = make_float4(col[0], col[1], col[2], col[3]);
        }
        __syncthreads();
        if (wmax <= lo) continue;  // this warp has no contributor in the batch
        for (int jj = min(cnt, wmax - lo) - 1; jj >= 0; jj -= RB) {
            float part[NT][RB];
            unsigned present = 0;  // bit u set when any lane of the warp has a valid sample of splat jj-u
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int j = jj - u;
                bool valid = (j >= 0) && ((lo + j) < last);
                float dx = 0.f, dy = 0.f, G = 0.f, a = 0.f;
                float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f);
                float2 g2 = make_float2(0.f, 0.f);
                if (valid) {
                    g1 = s_g1[j];
                    g2 = s_g2[j];
                    dx = g1.x - pxf; dy = g1.y - pyf;
                    const float power = -0.5f * (g1.z * dx * dx + g2.x * dy * dy) - g1.w * dx * dy;
                    G = __expf(power);
                    a = fminf(amax, g2.y * G);
                    valid = !(power > 0.0f) && (a >= ALPHA_MIN);
                }
                present |= (__ballot_sync(FULL, valid) != 0u) ? (1u << u) : 0u;
#pragma unroll
                for (int k = 0; k < NT; ++k) part[k][u] = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - a);
                    T *= ra;
                    const float fac = a * T;
                    const float4 col4 = s_col[j];
                    const float col[4] = {col4.x, col4.y, col4.z, col4.w};
                    float v_al = 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        part[6 + c][u] = fac * vo[c];
                        v_al += (col[c] * T - buf[c] * ra) * vo[c];
                        buf[c] += col[c] * fac;
                    }
                    v_al += tail * ra;
                    if (!GSPLAT || (g2.y * G <= 0.999f)) {
                        const float v_sigma = -g2.y * G * v_al;
                        const float gx = v_sigma * (g1.z * dx + g1.w * dy);
                        const float gy = v_sigma * (g1.w * dx + g2.x * dy);
                        part[0][u] = gx;
                        part[1][u] = gy;
                        part[2][u] = 0.5f * v_sigma * dx * dx;
                        part[3][u] = v_sigma * dx * dy;
                        part[4][u] = 0.5f * v_sigma * dy * dy;
                        part[5][u] = G * v_al;
                        if (ABS) {
                            part[6 + CH][u] = fabsf(gx);
                            part[(7 + CH) % NT][u] = fabsf(gy);
                        }
                    }
                }
            }
            if (present == 0u) continue;
            float tot[NT];
#pragma unroll
            for (int k = 0; k < NT; ++k) tot[k] = reduce_scatter<RB>(part[k], lane);
            const int j = jj - my_slot;
            if (writer && ((present >> my_slot) & 1u)) {
                const int g = s_id[j];
                atomicAdd(v_xy + 2 * g, tot[0] * sx);
                atomicAdd(v_xy + 2 * g + 1, tot[1] * sy);
                atomicAdd(v_conic + 3 * g, tot[2]);
                atomicAdd(v_conic + 3 * g + 1, tot[3]);
                atomicAdd(v_conic + 3 * g + 2, tot[4]);
                atomicAdd(v_opacity + g, tot[5]);
#pragma unroll
                for (int c = 0; c < CH; ++c) atomicAdd(v_colors + int64_t(g) * CH + c, tot[6 + c]);
                if (ABS) {
                    atomicAdd(v_xy_abs + 2 * g, tot[6 + CH]);
                    atomicAdd(v_xy_abs + 2 * g + 1, tot[(7 + CH) % NT]);
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

#ifndef B200GS_BWD_RB
#define B200GS_BWD_RB 8
#endif

template <int CH>
int bwd_dispatch(int mode, int width, int height, const int32_t* ranges, const int32_t* ids, const float* xy, const float* conic,
                 const float* opacity, const float* colors, const float* bg, const float* final_T, const int32_t* n_contrib,
                 const float* v_image, int64_t ps, int64_t cs, const float* v_alpha, float sx, float sy, float* v_xy, float* v_conic,
                 float* v_opacity, float* v_colors, float* v_xy_abs, cudaStream_t s) {
    const int gx = div_up(width, TILE), gy = div_up(height, TILE);
    dim3 grid(gx, gy);
    constexpr int RB = B200GS_BWD_RB;
#define B200GS_BWD_ARGS width, height, gx, (const int2*)ranges, ids, (const float2*)xy, conic, opacity, colors, bg, final_T, n_contrib, \
                        v_image, ps, cs, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs
    if (mode == B200GS_MODE_GSPLAT) {
        if (v_xy_abs)
            blend_bwd_kernel<CH, true, true, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
        else
            blend_bwd_kernel<CH, true, false, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
    } else {
        if (v_xy_abs)
            blend_bwd_kernel<CH, false, true, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
        else
            blend_bwd_kernel<CH, false, false, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
    }
#undef B200GS_BWD_ARGS
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
