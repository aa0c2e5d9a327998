// K6 / K7: per-tile front-to-back alpha compositing, forward and backward.
//
// Restates the published per-pixel loops of diff-gaussian-rasterization@59f5f77 renderCUDA (vanilla mode) and gsplat's
// rasterize_to_pixels (gsplat mode) — SURVEY.md §8c / Appendix B; the sources are not under /root/reference, the CPU
// restatement in oracle/gs_oracle.py::blend is the checker.
//
// One CTA per 16x16 tile, one thread per pixel; a warp owns an 8x4 pixel block.  Each round stages up to 256 splats of
// the tile's depth-sorted slab into shared memory (coalesced id read, L2-resident gathers of the 36 B splat record).
// While staging, the thread that fetched a splat also computes an 8-bit mask: which of the tile's eight 8x4 blocks
// the splat can reach with alpha >= 1/255 (exact convex minimum of the conic's quadratic over the block, same test as
// the tile culling of binning.cu).  Every warp then compacts the staged slab into its own index list with ballots, so
// its pixel loop only visits splats that can touch its block: ~half of the (warp, splat) visits of the plain loop
// disappear, and results stay bit-identical (a skipped splat would have failed the alpha test in all 32 lanes).
//
// Forward: the conic is pre-scaled by -0.5*log2(e) / -log2(e) while staging, so the per-(pixel,splat) body is
// 5 FP32 ops + one MUFU.EX2 + compare/blend; colours are fetched (one LDS.128) only by contributing lanes.
//
// Backward walks each warp's list in reverse, only up to the deepest contributor of the warp.  Each pixel evaluates RB
// consecutive list entries and keeps their 9 partial derivatives in registers; the warp then reduces them with a
// reduce-SCATTER butterfly (halving exchanges: RB -> RB/2 -> ... -> 1 value per lane) and one lane per splat issues
// the atomics: 32x fewer L2 atomics than the reference's per-pixel atomicAdd.
#include <stdlib.h>

#include "common.cuh"

namespace b200gs {

namespace {

constexpr int BLOCK_PIX = TILE * TILE;  // 256 threads
constexpr int NWARP = BLOCK_PIX / 32;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_STOP = 1e-4f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr unsigned FULL = 0xffffffffu;

// Element strides of the per-splat arrays: {2,3,1,CH} for separate contiguous arrays; {12,12,12,12} when all four
// pointers address columns of one [n,12] row buffer (the exchange format of the Gaussian-sharded renderer).
struct SplatStrides {
    int xs, cs, os, ks;
};

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

// Minimum of q(d) = (A dx^2 + C dy^2)/2 + B dx dy over the box [X0,X1]x[Y0,Y1] (d measured from the splat centre).
__device__ __forceinline__ float box_qmin(float A, float B, float C, float iA, float iC, float X0, float X1, float Y0, float Y1) {
    if (X0 <= 0.f && X1 >= 0.f && Y0 <= 0.f && Y1 >= 0.f) return 0.f;
    float qmin = 3.0e38f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float Xe = e ? X1 : X0;
        const float bx = B * Xe;
        const float ys = fminf(Y1, fmaxf(Y0, -bx * iC));
        qmin = fminf(qmin, fmaf(ys, fmaf(0.5f * C, ys, bx), 0.5f * A * Xe * Xe));
        const float Ye = e ? Y1 : Y0;
        const float by = B * Ye;
        const float xs = fminf(X1, fmaxf(X0, -by * iA));
        qmin = fminf(qmin, fmaf(xs, fmaf(0.5f * A, xs, by), 0.5f * C * Ye * Ye));
    }
    return qmin;
}

// bit w set <=> the splat can reach alpha >= 1/255 somewhere in warp w's 8x4 block of the tile at (ox, oy)
__device__ __forceinline__ unsigned block_mask(float mx, float my, float A, float B, float C, float opac, float ox, float oy) {
    const float o255 = 255.0f * opac;
    if (o255 <= 1.0f) return 0u;
    const float thresh = fmaf(__logf(o255), 1.0001f, 1e-3f);  // ln(255 o) + margin for fp32 / ex2.approx rounding
    const float iA = 1.0f / A, iC = 1.0f / C;
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
        const float X0 = ox + float((w & 1) << 3) - mx, Y0 = oy + float((w >> 1) << 2) - my;
        const float q = box_qmin(A, B, C, iA, iC, X0, X0 + 7.0f, Y0, Y0 + 3.0f);
        m |= (!(q > thresh)) ? (1u << w) : 0u;  // NaN -> keep
    }
    return m;
}

// Staged splats live in shared memory as 48-byte records (3 x float4): one base address per splat, immediate offsets.
//   forward : {x, y, a', b'} {c', opacity, col0, col1} {col2, col3, -, -}     (a',b',c' = conic pre-scaled by -0.5*log2e / -log2e)
//   backward: {x, y, A, B}   {C, opacity, col0, col1}  {col2, col3, id, -}
// Record DUMMY (index 256) has opacity 0: it fails the alpha test in every lane and pads the per-warp lists to a
// multiple of the unroll factor, so the pixel loops are straight-line code.
constexpr int DUMMY = BLOCK_PIX;
constexpr int LIST_PAD = 8;

// Every warp compacts the staged slab [0,top) into the ascending list of entries whose mask has its bit set.
// my_list[LIST_PAD + k] = k-th entry; the LIST_PAD slots in front and the slots after the end hold DUMMY.
template <int FRONT = LIST_PAD>
__device__ __forceinline__ int build_list(const unsigned char* __restrict__ s_mask, unsigned short* __restrict__ my_list, int top, int warp,
                                          unsigned lane) {
    int n = 0;
    if (lane < FRONT) my_list[lane] = (unsigned short)DUMMY;
    for (int c = 0; c < top; c += 32) {
        const int j = c + (int)lane;
        const bool hit = (j < top) && ((s_mask[j] >> warp) & 1u);
        const unsigned b = __ballot_sync(FULL, hit);
        if (hit) my_list[FRONT + n + __popc(b & ((1u << lane) - 1u))] = (unsigned short)j;
        n += __popc(b);
    }
    if (lane < LIST_PAD) my_list[FRONT + n + lane] = (unsigned short)DUMMY;
    __syncwarp();
    return n;
}

#ifndef B200GS_FWD_MINBLOCKS
#define B200GS_FWD_MINBLOCKS 1
#endif

template <int CH, bool GSPLAT>
__global__ void __launch_bounds__(BLOCK_PIX, B200GS_FWD_MINBLOCKS) blend_fwd_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const SplatStrides st, const float* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              float* __restrict__ image, int64_t pix_stride, int64_t ch_stride,
                                                              float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                                                              float* __restrict__ alpha_out) {
    __shared__ float4 s_rec[(BLOCK_PIX + 1) * 3];
    __shared__ unsigned char s_mask[BLOCK_PIX];
    __shared__ unsigned short s_list[NWARP][BLOCK_PIX + 2 * LIST_PAD];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31u;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float pxf = float(px) + off, pyf = float(py) + off;
    const float ox = float(blockIdx.x * TILE) + off, oy = float(blockIdx.y * TILE) + off;
    const float amax = GSPLAT ? 0.999f : 0.99f;
    if (tid < 3) s_rec[DUMMY * 3 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    bool done = !inside;
    float T = 1.0f;
    int last = 0;
    float C[4] = {0.f, 0.f, 0.f, 0.f};

    for (int base = 0; todo > 0; base += BLOCK_PIX, todo -= BLOCK_PIX) {
        if (__syncthreads_and(done)) break;
        const int cnt = min(BLOCK_PIX, todo);
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + base + tid);
            const float2 m = __ldg(reinterpret_cast<const float2*>(xy + int64_t(g) * st.xs));
            const float* cq = conic + int64_t(g) * st.cs;
            const float A = __ldg(cq), B = __ldg(cq + 1), Cc = __ldg(cq + 2);
            const float o = __ldg(opacity + int64_t(g) * st.os);
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * st.ks + c);
            s_rec[tid * 3 + 0] = make_float4(m.x, m.y, (-0.5f * LOG2E) * A, -LOG2E * B);
            s_rec[tid * 3 + 1] = make_float4((-0.5f * LOG2E) * Cc, o, col[0], col[1]);
            if (CH > 2) s_rec[tid * 3 + 2] = make_float4(col[2], col[3], 0.f, 0.f);
            s_mask[tid] = (unsigned char)block_mask(m.x, m.y, A, B, Cc, o, ox, oy);
        }
        __syncthreads();
        if (__all_sync(FULL, done)) continue;  // warp finished: only keeps the block barriers company
        const unsigned short* my_list = s_list[warp] + LIST_PAD;
        const int nl = build_list(s_mask, s_list[warp], cnt, warp, lane);
        for (int i0 = 0; i0 < nl; i0 += 4) {
            if (__all_sync(FULL, done)) break;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = my_list[i0 + u];                 // entries past nl are DUMMY (alpha 0)
                const float4 r0 = s_rec[j * 3 + 0];
                const float4 r1 = s_rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                // power * log2(e) = a' dx^2 + b' dx dy + c' dy^2
                const float p2 = fmaf(r1.x * dy, dy, fmaf(r0.w, dy, r0.z * dx) * dx);
                const float a = fminf(amax, r1.y * ex2_approx(p2));
                const float nT = fmaf(-a, T, T);
                const bool ok = !done && !(p2 > 0.0f) && !(a < ALPHA_MIN);
                const bool stop = ok && (GSPLAT ? (nT <= T_STOP) : (nT < T_STOP));
                const bool take = ok && !stop;
                const float w = take ? a * T : 0.f;
                C[0] = fmaf(r1.z, w, C[0]);
                if (CH > 1) C[1] = fmaf(r1.w, w, C[1]);
                if (CH > 2) {
                    const float2 r2 = *reinterpret_cast<const float2*>(&s_rec[j * 3 + 2]);
                    C[2] = fmaf(r2.x, w, C[2]);
                    if (CH > 3) C[3] = fmaf(r2.y, w, C[3]);
                }
                T = take ? nT : T;
                last = take ? base + j + 1 : last;
                done = done || stop;
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

// ---- forward with asynchronous slab staging -----------------------------------------------------------------------------
// Same arithmetic and the same per-warp lists as blend_fwd_kernel; what changes is how a batch reaches shared memory.
// The tile's slab is an indirection (sorted ids -> records scattered over the splat arrays), so every thread copies
// "its" splat of the NEXT batch with cp.async (LDGSTS: global -> shared without a register round trip, 8 B + 7 x 4 B from the
// separate arrays, or 3 x 16 B from a [n,12] row buffer) into the second of two staging buffers while the warps blend the current
// batch; the ids are prefetched two batches ahead in a register.  When a batch starts, its copies have landed long
// ago (cp.async.wait_all + barrier), each thread post-processes its own record in place (conic pre-scaling, 8-bit block
// mask) and the warps go on to the list compaction: the two dependent L2 round trips of the synchronous version
// (id -> record) are off the critical path of the tile.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// raw slot layout (floats): separate arrays -> {x, y, A, B | C, o, c0, c1 | c2, c3, -, -};  rows -> the 12 floats of the row
// {x, y, depth, A | B, C, comp, o | r, g, b, radius}
template <int CH, bool ROWS>
__device__ __forceinline__ void stage_async(float4* slot, int g, const SplatStrides& st, const float* __restrict__ xy,
                                            const float* __restrict__ conic, const float* __restrict__ opacity,
                                            const float* __restrict__ colors) {
    if (ROWS) {
        const float* row = xy + int64_t(g) * st.xs;     // xy points at column 0 of the row buffer
        cp_async16(slot, row);
        cp_async16(slot + 1, row + 4);
        cp_async16(slot + 2, row + 8);
    } else {
        float* f = reinterpret_cast<float*>(slot);
        cp_async8(f, xy + int64_t(g) * st.xs);
        const float* cq = conic + int64_t(g) * st.cs;
        cp_async4(f + 2, cq);
        cp_async4(f + 3, cq + 1);
        cp_async4(f + 4, cq + 2);
        cp_async4(f + 5, opacity + int64_t(g) * st.os);
#pragma unroll
        for (int c = 0; c < CH; ++c) cp_async4(f + 6 + c, colors + int64_t(g) * st.ks + c);
    }
}

// HITS: also mark every splat that contributed to at least one pixel (gsplat's `means2d.has_hit_any_pixels`, read by
// SelectiveAdam, optimizers.py:39, and exported as `acc_vis`, gsplat_v1_renderer.py:287): one byte store per (warp, contributing entry).
template <int CH, bool GSPLAT, bool ROWS, bool HITS>
__global__ void __launch_bounds__(BLOCK_PIX, B200GS_FWD_MINBLOCKS) blend_fwd_async_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const SplatStrides st, const float* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              float* __restrict__ image, int64_t pix_stride, int64_t ch_stride,
                                                              float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                                                              float* __restrict__ alpha_out, uint8_t* __restrict__ hit_any) {
    __shared__ float4 s_buf[2][(BLOCK_PIX + 1) * 3];
    __shared__ int32_t s_gid[HITS ? BLOCK_PIX + 1 : 1];
    __shared__ unsigned char s_mask[BLOCK_PIX];
    __shared__ unsigned short s_list[NWARP][BLOCK_PIX + 2 * LIST_PAD];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31u;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float pxf = float(px) + off, pyf = float(py) + off;
    const float ox = float(blockIdx.x * TILE) + off, oy = float(blockIdx.y * TILE) + off;
    const float amax = GSPLAT ? 0.999f : 0.99f;
    if (tid < 6) s_buf[tid / 3][DUMMY * 3 + tid % 3] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int2 range = ranges[tile];
    const int total = range.y - range.x;
    bool done = !inside;
    float T = 1.0f;
    int last = 0;
    float C[4] = {0.f, 0.f, 0.f, 0.f};

    // prologue: batch 0 in flight, ids of batch 1 in a register
    int cur_id = (tid < total) ? __ldg(ids + range.x + tid) : 0;
    if (tid < total) stage_async<CH, ROWS>(&s_buf[0][tid * 3], cur_id, st, xy, conic, opacity, colors);
    int next_id = (BLOCK_PIX + tid < total) ? __ldg(ids + range.x + BLOCK_PIX + tid) : 0;

    int buf = 0;
    for (int base = 0; base < total; base += BLOCK_PIX, buf ^= 1) {
        cp_async_wait_all();
        if (__syncthreads_and(done)) break;          // also: every thread's copies of this batch are visible
        const int cnt = min(BLOCK_PIX, total - base);
        float4* s_rec = s_buf[buf];
        if (tid < cnt) {
            const float4 q0 = s_rec[tid * 3 + 0], q1 = s_rec[tid * 3 + 1], q2 = s_rec[tid * 3 + 2];
            float mx, my, A, B, Cc, o, c0, c1, c2, c3;
            if (ROWS) { mx = q0.x; my = q0.y; A = q0.w; B = q1.x; Cc = q1.y; o = q1.w; c0 = q2.x; c1 = q2.y; c2 = q2.z; c3 = 0.f; }
            else      { mx = q0.x; my = q0.y; A = q0.z; B = q0.w; Cc = q1.x; o = q1.y; c0 = q1.z; c1 = q1.w; c2 = q2.x; c3 = q2.y; }
            s_rec[tid * 3 + 0] = make_float4(mx, my, (-0.5f * LOG2E) * A, -LOG2E * B);
            s_rec[tid * 3 + 1] = make_float4((-0.5f * LOG2E) * Cc, o, c0, CH > 1 ? c1 : 0.f);
            if (CH > 2) s_rec[tid * 3 + 2] = make_float4(c2, CH > 3 ? c3 : 0.f, 0.f, 0.f);
            s_mask[tid] = (unsigned char)block_mask(mx, my, A, B, Cc, o, ox, oy);
            if (HITS) s_gid[tid] = cur_id;
        }
        cur_id = next_id;
        // next batch: the other buffer was last read by the blend loop of the previous iteration, which every warp left
        // before the barrier above
        if (base + BLOCK_PIX + tid < total) stage_async<CH, ROWS>(&s_buf[buf ^ 1][tid * 3], next_id, st, xy, conic, opacity, colors);
        next_id = (base + 2 * BLOCK_PIX + tid < total) ? __ldg(ids + range.x + base + 2 * BLOCK_PIX + tid) : 0;
        __syncthreads();
        if (__all_sync(FULL, done)) continue;  // warp finished: only keeps the block barriers company
        const unsigned short* my_list = s_list[warp] + LIST_PAD;
        const int nl = build_list(s_mask, s_list[warp], cnt, warp, lane);
        for (int i0 = 0; i0 < nl; i0 += 4) {
            if (__all_sync(FULL, done)) break;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = my_list[i0 + u];                 // entries past nl are DUMMY (alpha 0)
                const float4 r0 = s_rec[j * 3 + 0];
                const float4 r1 = s_rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                const float p2 = fmaf(r1.x * dy, dy, fmaf(r0.w, dy, r0.z * dx) * dx);
                const float a = fminf(amax, r1.y * ex2_approx(p2));
                const float nT = fmaf(-a, T, T);
                const bool ok = !done && !(p2 > 0.0f) && !(a < ALPHA_MIN);
                const bool stop = ok && (GSPLAT ? (nT <= T_STOP) : (nT < T_STOP));
                const bool take = ok && !stop;
                const float w = take ? a * T : 0.f;
                C[0] = fmaf(r1.z, w, C[0]);
                if (CH > 1) C[1] = fmaf(r1.w, w, C[1]);
                if (CH > 2) {
                    const float2 r2 = *reinterpret_cast<const float2*>(&s_rec[j * 3 + 2]);
                    C[2] = fmaf(r2.x, w, C[2]);
                    if (CH > 3) C[3] = fmaf(r2.y, w, C[3]);
                }
                T = take ? nT : T;
                last = take ? base + j + 1 : last;
                done = done || stop;
                if (HITS) {
                    if (__any_sync(FULL, take) && lane == 0) hit_any[s_gid[j]] = 1;
                }
            }
        }
    }
    cp_async_wait_all();
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

// VS ("value scatter", experimental, opt-in with B200GS_BWD_VS=1, CH == 3, RB == 4, no absgrad): after the two entry-splitting
// steps of the reduce-scatter the nine totals of an entry are reduced over the remaining 8 lanes with a reduce-scatter over
// the VALUES (5 + 3 + 3 shuffles instead of 27), so each of the 8 lanes ends up owning one total (lane 0 of the group owns
// the coupled pair sum(vs dx), sum(vs dy)) and issues its own atomic: two lane-parallel REDs instead of nine serial ones.
template <int CH, bool GSPLAT, bool ABS, int RB, bool VS = false>
__global__ void __launch_bounds__(BLOCK_PIX, (RB == 4 && !ABS) ? 4 : 1) blend_bwd_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                              const int32_t* __restrict__ ids, const SplatStrides st, const float* __restrict__ xy,
                                                              const float* __restrict__ conic, const float* __restrict__ opacity,
                                                              const float* __restrict__ colors, const float* __restrict__ bg,
                                                              const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
                                                              const float* __restrict__ v_image, int64_t pix_stride, int64_t ch_stride,
                                                              const float* __restrict__ v_alpha, float sx, float sy,
                                                              float* __restrict__ v_xy, float* __restrict__ v_conic,
                                                              float* __restrict__ v_opacity, float* __restrict__ v_colors,
                                                              float* __restrict__ v_xy_abs) {
    // reduced per splat: the six moments  sum(go), sum(vs dx), sum(vs dy), sum(vs dx^2), sum(vs dx dy), sum(vs dy^2)  (vs = dL/dsigma,
    // go = dL/dopacity) — the conic/mean gradients are linear in them, so the writer lane finishes the products once per
    // splat instead of every lane per sample — plus the colour sums, plus |grad xy| when the absgrad side channel is on.
    constexpr int NT = 6 + CH + (ABS ? 2 : 0);
    static_assert(!VS || (CH == 3 && RB == 4 && !ABS), "value-scatter variant: 3 channels, RB 4, no absgrad");
    __shared__ float4 s_rec[(BLOCK_PIX + 1) * 3];
    __shared__ unsigned char s_mask[BLOCK_PIX];
    __shared__ unsigned short s_list[NWARP][BLOCK_PIX + 2 * LIST_PAD];
    __shared__ int s_wmax[NWARP];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31u;
    // VS: which total lane (lane & 7) of an 8-lane group owns, and where it goes
    float* vs_base = nullptr;
    int vs_stride = 0;
    float vs_scale = 1.0f;
    if (VS) {
        switch (lane & 7u) {
            case 0: vs_base = v_xy; vs_stride = st.xs; vs_scale = sx; break;              // + the y component, see the writer
            case 1: vs_base = v_opacity; vs_stride = st.os; break;
            case 2: vs_base = v_conic; vs_stride = st.cs; vs_scale = 0.5f; break;
            case 3: vs_base = v_conic + 1; vs_stride = st.cs; break;
            case 4: vs_base = v_conic + 2; vs_stride = st.cs; vs_scale = 0.5f; break;
            case 5: vs_base = v_colors; vs_stride = st.ks; break;
            case 6: vs_base = v_colors + 1; vs_stride = st.ks; break;
            default: vs_base = v_colors + 2; vs_stride = st.ks; break;
        }
    }
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float pxf = float(px) + off, pyf = float(py) + off;
    const float ox = float(blockIdx.x * TILE) + off, oy = float(blockIdx.y * TILE) + off;
    const float amax = GSPLAT ? 0.999f : 0.99f;
    const int64_t pix = int64_t(py) * width + px;
    if (tid < 3) s_rec[DUMMY * 3 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int2 range = ranges[tile];
    const float Tf = inside ? final_T[pix] : 0.f;
    const int last = inside ? n_contrib[pix] : 0;
    float vo[4] = {0.f, 0.f, 0.f, 0.f};
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        vo[c] = inside ? __ldg(v_image + pix * pix_stride + c * ch_stride) : 0.f;
        if (bg) bg_dot += __ldg(bg + c) * vo[c];
    }
    const float va = (v_alpha && inside) ? __ldg(v_alpha + pix) : 0.f;
    const float tail = Tf * (va - bg_dot);  // d(out)/d(alpha_i) through everything behind the last contributor

    const int wmax = __reduce_max_sync(FULL, last);
    if (lane == 0) s_wmax[warp] = wmax;
    __syncthreads();
    int max_last = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) max_last = max(max_last, s_wmax[w]);
    if (max_last == 0) return;

    float T = Tf;
    float D = 0.f;   // <colour accumulated behind the current splat, v_image> for this pixel
    const int my_slot = rs_slot<RB>(lane);
    const bool writer = (lane & (32 / RB - 1)) == 0;

    for (int hi = max_last; hi > 0; hi -= BLOCK_PIX) {
        const int lo = max(0, hi - BLOCK_PIX);
        const int cnt = hi - lo;
        __syncthreads();
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + lo + tid);
            const float2 m = __ldg(reinterpret_cast<const float2*>(xy + int64_t(g) * st.xs));
            const float* cq = conic + int64_t(g) * st.cs;
            const float A = __ldg(cq), B = __ldg(cq + 1), Cc = __ldg(cq + 2);
            const float o = __ldg(opacity + int64_t(g) * st.os);
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * st.ks + c);
            s_rec[tid * 3 + 0] = make_float4(m.x, m.y, (-0.5f * LOG2E) * A, -LOG2E * B);
            s_rec[tid * 3 + 1] = make_float4((-0.5f * LOG2E) * Cc, o, col[0], col[1]);
            s_rec[tid * 3 + 2] = make_float4(col[2], col[3], __int_as_float(g), 0.f);
            s_mask[tid] = (unsigned char)block_mask(m.x, m.y, A, B, Cc, o, ox, oy);
        }
        __syncthreads();
        if (wmax <= lo) continue;  // this warp has no contributor in the batch
        const unsigned short* my_list = s_list[warp] + LIST_PAD;   // my_list[-LIST_PAD..-1] are DUMMY
        const int nl = build_list(s_mask, s_list[warp], min(cnt, wmax - lo), warp, lane);
        for (int ii = nl - 1; ii >= 0; ii -= RB) {
            float part[NT][RB];
            unsigned present = 0;  // bit u set when any lane of the warp has a valid sample of list entry ii-u
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int j = my_list[ii - u];
                const float4 r0 = s_rec[j * 3 + 0];
                const float4 r1 = s_rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                // same arithmetic as the forward: power * log2(e) = a' dx^2 + b' dx dy + c' dy^2
                const float p2 = fmaf(r1.x * dy, dy, fmaf(r0.w, dy, r0.z * dx) * dx);
                const float G = ex2_approx(p2);
                const float a = fminf(amax, r1.y * G);
                const bool valid = ((lo + j) < last) && !(p2 > 0.0f) && (a >= ALPHA_MIN);
                present |= (__ballot_sync(FULL, valid) != 0u) ? (1u << u) : 0u;
                float vs = 0.f, fac = 0.f, go = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - a);
                    T *= ra;
                    fac = a * T;
                    float S = r1.z * vo[0];
                    if (CH > 1) S = fmaf(r1.w, vo[1], S);
                    if (CH > 2) {
                        const float2 r2 = *reinterpret_cast<const float2*>(&s_rec[j * 3 + 2]);
                        S = fmaf(r2.x, vo[2], S);
                        if (CH > 3) S = fmaf(r2.y, vo[3], S);
                    }
                    // dL/dalpha = T <c, v> - (<colour behind, v> + T_final (bg.v - v_alpha)) / (1 - alpha);  D = <colour behind, v>
                    const float v_al = fmaf(T, S, ra * (tail - D));
                    D = fmaf(fac, S, D);
                    if (!GSPLAT || (r1.y * G <= 0.999f)) {
                        go = G * v_al;
                        vs = -r1.y * go;
                    }
                }
                const float t1 = vs * dx, t2 = vs * dy;
                part[0][u] = go;
                part[1][u] = t1;
                part[2][u] = t2;
                part[3][u] = t1 * dx;
                part[4][u] = t1 * dy;
                part[5][u] = t2 * dy;
#pragma unroll
                for (int c = 0; c < CH; ++c) part[6 + c][u] = fac * vo[c];
                if (ABS) {
                    part[6 + CH][u] = fabsf(r0.z * t1 + 0.5f * r0.w * t2) * (2.0f / LOG2E);
                    part[(7 + CH) % NT][u] = fabsf(0.5f * r0.w * t1 + r1.x * t2) * (2.0f / LOG2E);
                }
            }
            if (present == 0u) continue;
            if (VS) {
                const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
                float sv[NT];      // entry split: this lane's entry (slot my_slot), summed over the lanes that differ in bits 16 and 8
#pragma unroll
                for (int k = 0; k < NT; ++k)
                    sv[k] = rs_step(rs_step(part[k][0], part[k][2], b16, 16), rs_step(part[k][1], part[k][3], b16, 16), b8, 8);
                // value split over the 8 lanes of the group.  sv: 0 go, 1 t1, 2 t2, 3..5 conic moments, 6..8 colours
                const float r0 = rs_step(sv[1], sv[5], b4, 4), r1 = rs_step(sv[2], sv[6], b4, 4);
                const float r3 = rs_step(sv[3], sv[7], b4, 4), r4 = rs_step(sv[4], sv[8], b4, 4);
                const float r2 = sv[0] + __shfl_xor_sync(FULL, sv[0], 4);
                const float u0 = rs_step(r0, r3, b2, 2), u1 = rs_step(r1, r4, b2, 2);
                const float u2 = r2 + __shfl_xor_sync(FULL, r2, 2);
                const float w = rs_step(u0, u1, b1, 1);                     // lane&7: 0 t1, 2 m3, 3 m4, 4 m5, 5 c0, 6 c1, 7 c2
                const float t2tot = u1 + __shfl_xor_sync(FULL, u1, 1);      // valid on lanes 0/1 of the group
                const float gotot = u2 + __shfl_xor_sync(FULL, u2, 1);      // valid on lanes 0/1 of the group
                if ((present >> my_slot) & 1u) {
                    const int j = my_list[ii - my_slot];
                    const int g = __float_as_int(s_rec[j * 3 + 2].z);
                    const unsigned l8 = lane & 7u;
                    float outv = vs_scale * (l8 == 1u ? gotot : w);
                    float outy = 0.f;
                    if (l8 == 0u) {
                        const float4 r0c = s_rec[j * 3 + 0];
                        const float A = r0c.z * (-2.0f / LOG2E), B = r0c.w * (-1.0f / LOG2E), Cc = s_rec[j * 3 + 1].x * (-2.0f / LOG2E);
                        outv = (A * w + B * t2tot) * sx;
                        outy = (B * w + Cc * t2tot) * sy;
                    }
                    float* dst = vs_base + int64_t(g) * vs_stride;
                    atomicAdd(dst, outv);
                    if (l8 == 0u) atomicAdd(dst + 1, outy);
                }
                continue;
            }
            float tot[NT];
#pragma unroll
            for (int k = 0; k < NT; ++k) tot[k] = reduce_scatter<RB>(part[k], lane);
            if (writer && ((present >> my_slot) & 1u)) {
                const int j = my_list[ii - my_slot];
                const float4 r0 = s_rec[j * 3 + 0];
                const float A = r0.z * (-2.0f / LOG2E), B = r0.w * (-1.0f / LOG2E), Cc = s_rec[j * 3 + 1].x * (-2.0f / LOG2E);
                const int g = __float_as_int(s_rec[j * 3 + 2].z);
                float* vx = v_xy + int64_t(g) * st.xs;
                float* vc = v_conic + int64_t(g) * st.cs;
                atomicAdd(vx, (A * tot[1] + B * tot[2]) * sx);
                atomicAdd(vx + 1, (B * tot[1] + Cc * tot[2]) * sy);
                atomicAdd(vc, 0.5f * tot[3]);
                atomicAdd(vc + 1, tot[4]);
                atomicAdd(vc + 2, 0.5f * tot[5]);
                atomicAdd(v_opacity + int64_t(g) * st.os, tot[0]);
#pragma unroll
                for (int c = 0; c < CH; ++c) atomicAdd(v_colors + int64_t(g) * st.ks + c, tot[6 + c]);
                if (ABS) {
                    atomicAdd(v_xy_abs + 2 * g, tot[6 + CH]);
                    atomicAdd(v_xy_abs + 2 * g + 1, tot[(7 + CH) % NT]);
                }
            }
        }
    }
}

// ---- backward, transpose-reduce variant (default) -----------------------------------------------------------------------
// The butterfly above spends 63 of its 119 warp instructions per (warp, list entry) on moving nine partial sums between
// lanes (SHFL issues at one warp instruction per clock per SM) and on a single writer lane.  Here each pixel lane only
// produces TWO numbers per entry —  go = dL/d(opacity-weighted Gaussian)  and  fac = alpha*T  — and stores them as one
// float2 into a per-warp shared-memory tile [GE entries][32 lanes].  After GE = 16 entries the roles flip: lane (e, h)
// owns entry e and the 16 pixels of half h of the warp's 8x4 block, reads its row with 128-bit loads (rows are padded by
// 16 B: conflict-free in both directions) and accumulates, in registers and with compile-time pixel coordinates as FFMA
// immediates, the six coordinate moments of go (the mean/conic gradients are linear in them) and the colour sums
// sum(fac * v_image[pixel]) (v_image of the warp's 32 pixels sits in shared memory, read as broadcasts).  One xor-16
// exchange combines the halves; then lane (e, 0) finishes the mean/conic algebra of entry e and issues its 5 atomics
// while lane (e, 1) issues opacity + colours: 16 entries are written by 32 lanes in parallel.  No shuffles in the
// reduction, no serial writer: ~12 instead of 63 instructions per (warp, entry) after the evaluation.
// With the absgrad side channel (gsplat's `means2d.absgrad`, vanilla_density_controller.py:112-113) each pixel stores two more
// values per entry, |dL/dmean2D.x| and |dL/dmean2D.y| of its own sample (they are summed without signs, so they cannot be
// derived from the moments): float4 rows, 3 CTAs per SM instead of 4.
constexpr int GE = 16;                          // list entries per reduction group
constexpr int BWD_LIST = BLOCK_PIX + GE + 8;    // per-warp list: GE dummies in front (reverse walk), 8 behind

template <bool ABS>
struct TrSmem {
    static constexpr int VAL_BYTES = ABS ? 16 : 8;              // per (entry, pixel): {go, fac} or {go, fac, |gx|, |gy|}
    static constexpr int VAL_ROW = 32 * VAL_BYTES + 16;         // bytes per entry row of the value tile (+ pad)
    float4 rec[(BLOCK_PIX + 1) * 3];
    float4 vo[NWARP][32];
    unsigned char val[NWARP][GE * VAL_ROW];
    unsigned short list[NWARP][BWD_LIST];
    unsigned char mask[BLOCK_PIX];
    int wmax[NWARP];
};
static_assert(sizeof(TrSmem<false>) <= 57344, "4 CTAs per SM need <= 56 KB each");

template <int CH, bool GSPLAT, bool ABS>
__global__ void __launch_bounds__(BLOCK_PIX, ABS ? 3 : 4) blend_bwd_tr_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                                    const int32_t* __restrict__ ids, const SplatStrides st,
                                                                    const float* __restrict__ xy, const float* __restrict__ conic,
                                                                    const float* __restrict__ opacity, const float* __restrict__ colors,
                                                                    const float* __restrict__ bg, const float* __restrict__ final_T,
                                                                    const int32_t* __restrict__ n_contrib, const float* __restrict__ v_image,
                                                                    int64_t pix_stride, int64_t ch_stride, const float* __restrict__ v_alpha,
                                                                    float sx, float sy, float* __restrict__ v_xy, float* __restrict__ v_conic,
                                                                    float* __restrict__ v_opacity, float* __restrict__ v_colors,
                                                                    float* __restrict__ v_xy_abs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef TrSmem<ABS> Smem;
    constexpr int VAL_ROW = Smem::VAL_ROW, VAL_BYTES = Smem::VAL_BYTES;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    float4* s_rec = sm.rec;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31u;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float pxf = float(px) + off, pyf = float(py) + off;
    const float ox = float(blockIdx.x * TILE) + off, oy = float(blockIdx.y * TILE) + off;
    const float amax = GSPLAT ? 0.999f : 0.99f;
    const int64_t pix = int64_t(py) * width + px;
    if (tid < 3) s_rec[DUMMY * 3 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int2 range = ranges[tile];
    const float Tf = inside ? final_T[pix] : 0.f;
    const int last = inside ? n_contrib[pix] : 0;
    float vo[4] = {0.f, 0.f, 0.f, 0.f};
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        vo[c] = inside ? __ldg(v_image + pix * pix_stride + c * ch_stride) : 0.f;
        if (bg) bg_dot += __ldg(bg + c) * vo[c];
    }
    sm.vo[warp][lane] = make_float4(vo[0], vo[1], vo[2], vo[3]);
    const float va = (v_alpha && inside) ? __ldg(v_alpha + pix) : 0.f;
    const float tail = Tf * (va - bg_dot);  // d(out)/d(alpha_i) through everything behind the last contributor

    const int wmax = __reduce_max_sync(FULL, last);
    if (lane == 0) sm.wmax[warp] = wmax;
    __syncthreads();
    int max_last = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) max_last = max(max_last, sm.wmax[w]);
    if (max_last == 0) return;

    float T = Tf;
    float D = 0.f;   // <colour accumulated behind the current splat, v_image> for this pixel
    unsigned char* my_val = sm.val[warp] + lane * VAL_BYTES;      // eval phase: this pixel's value column
    const int re = lane & (GE - 1), rh = lane >> 4;               // reduce phase: entry and pixel half of this lane
    const float4* my_row = reinterpret_cast<const float4*>(sm.val[warp] + re * VAL_ROW + rh * 16 * VAL_BYTES);
    const float4* my_vo = sm.vo[warp] + rh * 16;
    // origin of the warp's 8x4 block of pixel samples; the reduce lane's pixels are rows 2 rh, 2 rh + 1 of it
    const float bx0 = ox + float((warp & 1) << 3), by0 = oy + float((warp >> 1) << 2);
    const float hh = float(2 * rh);

    for (int hi = max_last; hi > 0; hi -= BLOCK_PIX) {
        const int lo = max(0, hi - BLOCK_PIX);
        const int cnt = hi - lo;
        __syncthreads();
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + lo + tid);
            const float2 m = __ldg(reinterpret_cast<const float2*>(xy + int64_t(g) * st.xs));
            const float* cq = conic + int64_t(g) * st.cs;
            const float A = __ldg(cq), B = __ldg(cq + 1), Cc = __ldg(cq + 2);
            const float o = __ldg(opacity + int64_t(g) * st.os);
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * st.ks + c);
            s_rec[tid * 3 + 0] = make_float4(m.x, m.y, (-0.5f * LOG2E) * A, -LOG2E * B);
            s_rec[tid * 3 + 1] = make_float4((-0.5f * LOG2E) * Cc, o, col[0], col[1]);
            s_rec[tid * 3 + 2] = make_float4(col[2], col[3], __int_as_float(g), 0.f);
            sm.mask[tid] = (unsigned char)block_mask(m.x, m.y, A, B, Cc, o, ox, oy);
        }
        __syncthreads();
        if (wmax <= lo) continue;  // this warp has no contributor in the batch
        const unsigned short* my_list = sm.list[warp] + GE;   // my_list[-GE..-1] are DUMMY
        const int nl = build_list<GE>(sm.mask, sm.list[warp], min(cnt, wmax - lo), warp, lane);
        const int rel_last = last - lo;                        // entry j of this batch is in front of the pixel's last contributor iff j < rel_last
        for (int ii = nl - 1; ii >= 0; ii -= GE) {
            unsigned vm = 0;   // bit u: this pixel has a valid sample of list entry ii-u
#pragma unroll
            for (int u = 0; u < GE; ++u) {
                const int j = my_list[ii - u];
                const float4 r0 = s_rec[j * 3 + 0];
                const float4 r1 = s_rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                // same arithmetic as the forward: power * log2(e) = a' dx^2 + b' dx dy + c' dy^2
                const float p2 = fmaf(r1.x * dy, dy, fmaf(r0.w, dy, r0.z * dx) * dx);
                const float G = ex2_approx(p2);
                const float oG = r1.y * G;
                const float a = fminf(amax, oG);
                const bool valid = (j < rel_last) && !(p2 > 0.0f) && (a >= ALPHA_MIN);
                float go = 0.f, fac = 0.f, gax = 0.f, gay = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - a);
                    T *= ra;
                    fac = a * T;
                    float S = r1.z * vo[0];
                    if (CH > 1) S = fmaf(r1.w, vo[1], S);
                    if (CH > 2) {
                        const float2 r2 = *reinterpret_cast<const float2*>(&s_rec[j * 3 + 2]);
                        S = fmaf(r2.x, vo[2], S);
                        if (CH > 3) S = fmaf(r2.y, vo[3], S);
                    }
                    // dL/dalpha = T <c, v> - (<colour behind, v> + T_final (bg.v - v_alpha)) / (1 - alpha);  D = <colour behind, v>
                    const float v_al = fmaf(T, S, ra * (tail - D));
                    D = fmaf(fac, S, D);
                    if (!GSPLAT || (oG <= 0.999f)) go = G * v_al;
                    vm |= 1u << u;
                    if (ABS) {   // |dL/dmean2D| of this sample: |v_sigma (A dx + B dy)|, |v_sigma (B dx + C dy)|, v_sigma = -opacity * go
                        const float vsg = r1.y * go * (2.0f / LOG2E);
                        gax = fabsf(vsg * fmaf(r0.z, dx, 0.5f * r0.w * dy));
                        gay = fabsf(vsg * fmaf(0.5f * r0.w, dx, r1.x * dy));
                    }
                }
                if (ABS) *reinterpret_cast<float4*>(my_val + u * VAL_ROW) = make_float4(go, fac, gax, gay);
                else *reinterpret_cast<float2*>(my_val + u * VAL_ROW) = make_float2(go, fac);
            }
            const unsigned present = __reduce_or_sync(FULL, vm);
            __syncwarp();   // the tile stores of all lanes are visible to the row loads below
            if (present != 0u) {
                // ---- reduce: lane (re, rh) sums entry ii-re over pixels 16 rh .. 16 rh + 15 (local x = k & 7, local row = k >> 3)
                float R0 = 0.f, R0x = 0.f, R0xx = 0.f, R1 = 0.f, R1x = 0.f, R1xx = 0.f;
                float Cs[4] = {0.f, 0.f, 0.f, 0.f};
                float Ax = 0.f, Ay = 0.f;
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    float4 v;                             // {go, fac} of pixels 2 k2, 2 k2 + 1
                    if (ABS) {
                        const float4 va = my_row[2 * k2], vb = my_row[2 * k2 + 1];
                        v = make_float4(va.x, va.y, vb.x, vb.y);
                        Ax += va.z + vb.z;
                        Ay += va.w + vb.w;
                    } else {
                        v = my_row[k2];
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = 2 * k2 + e;
                        const float gk = e ? v.z : v.x, fk = e ? v.w : v.y;
                        const float x = float(k & 7);
                        if (k < 8) { R0 += gk; if (k & 7) { R0x = fmaf(gk, x, R0x); R0xx = fmaf(gk, x * x, R0xx); } }
                        else       { R1 += gk; if (k & 7) { R1x = fmaf(gk, x, R1x); R1xx = fmaf(gk, x * x, R1xx); } }
                        const float4 w = my_vo[k];
                        Cs[0] = fmaf(fk, w.x, Cs[0]);
                        if (CH > 1) Cs[1] = fmaf(fk, w.y, Cs[1]);
                        if (CH > 2) Cs[2] = fmaf(fk, w.z, Cs[2]);
                        if (CH > 3) Cs[3] = fmaf(fk, w.w, Cs[3]);
                    }
                }
                // moments in block coordinates (x = 0..7, y = 2 rh + {0, 1}), then both halves
                float S0 = R0 + R1, Sx = R0x + R1x, Sxx = R0xx + R1xx;
                float Sy = fmaf(hh, S0, R1), Sxy = fmaf(hh, Sx, R1x), Syy = fmaf(hh * hh, S0, fmaf(2.0f * hh, R1, R1));
                S0 += __shfl_xor_sync(FULL, S0, 16);
                Sx += __shfl_xor_sync(FULL, Sx, 16);
                Sy += __shfl_xor_sync(FULL, Sy, 16);
                Sxx += __shfl_xor_sync(FULL, Sxx, 16);
                Sxy += __shfl_xor_sync(FULL, Sxy, 16);
                Syy += __shfl_xor_sync(FULL, Syy, 16);
#pragma unroll
                for (int c = 0; c < CH; ++c) Cs[c] += __shfl_xor_sync(FULL, Cs[c], 16);
                if (ABS) {
                    Ax += __shfl_xor_sync(FULL, Ax, 16);
                    Ay += __shfl_xor_sync(FULL, Ay, 16);
                }
                if ((present >> re) & 1u) {
                    const int j = my_list[ii - re];
                    const int g = __float_as_int(s_rec[j * 3 + 2].z);
                    if (rh == 0) {
                        const float4 r0 = s_rec[j * 3 + 0];
                        const float4 r1 = s_rec[j * 3 + 1];
                        const float A = r0.z * (-2.0f / LOG2E), B = r0.w * (-1.0f / LOG2E), Cc = r1.x * (-2.0f / LOG2E);
                        const float ex = r0.x - bx0, ey = r0.y - by0;   // dx = ex - x, dy = ey - y
                        const float no = -r1.y;                        // dL/dsigma = -opacity * go
                        const float M1 = no * fmaf(ex, S0, -Sx), M2 = no * fmaf(ey, S0, -Sy);
                        const float M3 = no * fmaf(ex, fmaf(ex, S0, -2.0f * Sx), Sxx);
                        const float M4 = no * (fmaf(ex, fmaf(ey, S0, -Sy), Sxy) - ey * Sx);
                        const float M5 = no * fmaf(ey, fmaf(ey, S0, -2.0f * Sy), Syy);
                        float* vx = v_xy + int64_t(g) * st.xs;
                        float* vc = v_conic + int64_t(g) * st.cs;
                        atomicAdd(vx, (A * M1 + B * M2) * sx);
                        atomicAdd(vx + 1, (B * M1 + Cc * M2) * sy);
                        atomicAdd(vc, 0.5f * M3);
                        atomicAdd(vc + 1, M4);
                        atomicAdd(vc + 2, 0.5f * M5);
                    } else {
                        atomicAdd(v_opacity + int64_t(g) * st.os, S0);
#pragma unroll
                        for (int c = 0; c < CH; ++c) atomicAdd(v_colors + int64_t(g) * st.ks + c, Cs[c]);
                        if (ABS) {
                            atomicAdd(v_xy_abs + 2 * int64_t(g), Ax);
                            atomicAdd(v_xy_abs + 2 * int64_t(g) + 1, Ay);
                        }
                    }
                }
            }
            __syncwarp();   // the rows are read: the next group may overwrite the tile
        }
    }
}

// ---- backward, tensor-core reduction variant -------------------------------------------------------------------------
// The cross-lane reduction of the backward IS a small dense contraction: for the 8 list entries of a group, every
// output is  sum_over_the_32_pixel_lanes( weight[row][lane] * value[lane][entry] )  with weights that are FIXED for the warp:
//   rows 0..5  : 1, cx, cy, cx^2, cx*cy, cy^2   (pixel coordinates relative to the centre of the warp's 8x4 block)   x  g = dL/d(opacity)
//   rows 8..8+CH-1 : v_image[channel] of the lane's pixel                                                             x  f = alpha*T
// (the mean/conic gradients are linear in the six pixel-coordinate moments of g, see the writer below).  That is a
// [16 x 64] x [64 x 8] GEMM per group: mma.sync.m16n8k8 TF32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi; the coordinate
// weights are exact in TF32) -> fp32-level accuracy, ~23 issue slots per splat for reduction + writer instead of ~58 with
// the shuffle butterfly.  Values go lane -> fragment through a 2 KB per-warp shared-memory tile (XOR-swizzled: both the
// stores and the fragment loads are bank-conflict free).
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CH, bool GSPLAT>
__global__ void __launch_bounds__(BLOCK_PIX, 3) blend_bwd_mma_kernel(int width, int height, int grid_x, const int2* __restrict__ ranges,
                                                                     const int32_t* __restrict__ ids, const SplatStrides st,
                                                                     const float* __restrict__ xy, const float* __restrict__ conic,
                                                                     const float* __restrict__ opacity, const float* __restrict__ colors,
                                                                     const float* __restrict__ bg, const float* __restrict__ final_T,
                                                                     const int32_t* __restrict__ n_contrib, const float* __restrict__ v_image,
                                                                     int64_t pix_stride, int64_t ch_stride, const float* __restrict__ v_alpha,
                                                                     float sx, float sy, float* __restrict__ v_xy, float* __restrict__ v_conic,
                                                                     float* __restrict__ v_opacity, float* __restrict__ v_colors) {
    constexpr int GB = 8;  // list entries per group = N of the MMA tile
    __shared__ float4 s_rec[(BLOCK_PIX + 1) * 3];
    __shared__ unsigned char s_mask[BLOCK_PIX];
    __shared__ unsigned short s_list[NWARP][BLOCK_PIX + 2 * LIST_PAD];
    __shared__ float s_b[NWARP][2][GB][32];
    __shared__ int s_wmax[NWARP];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const unsigned lane = tid & 31u;
    const int gid = lane >> 2, tig = lane & 3;
    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = (px < width) && (py < height);
    const float off = GSPLAT ? 0.5f : 0.0f;
    const float pxf = float(px) + off, pyf = float(py) + off;
    const float ox = float(blockIdx.x * TILE) + off, oy = float(blockIdx.y * TILE) + off;
    // centre of this warp's 8x4 block of pixel samples
    const float bcx = ox + float((warp & 1) << 3) + 3.5f, bcy = oy + float((warp >> 1) << 2) + 1.5f;
    const float amax = GSPLAT ? 0.999f : 0.99f;
    const int64_t pix = int64_t(py) * width + px;
    if (tid < 3) s_rec[DUMMY * 3 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int2 range = ranges[tile];
    const float Tf = inside ? final_T[pix] : 0.f;
    const int last = inside ? n_contrib[pix] : 0;
    float vo[4] = {0.f, 0.f, 0.f, 0.f};
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        vo[c] = inside ? __ldg(v_image + pix * pix_stride + c * ch_stride) : 0.f;
        if (bg) bg_dot += __ldg(bg + c) * vo[c];
    }
    const float va = (v_alpha && inside) ? __ldg(v_alpha + pix) : 0.f;
    const float tail = Tf * (va - bg_dot);

    // fixed A fragments.  G part (k-steps 0..3): rows 0..5 in a0/a2;  F part (k-steps 4..7): rows 8..8+CH-1 in a1/a3 (hi + lo)
    uint32_t aG0[4], aG2[4], aF1h[4], aF1l[4], aF3h[4], aF3l[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 8 * s4 + tig + 4 * h;                 // pixel lane this fragment element multiplies
            const float cx = float(k & 7) - 3.5f, cy = float(k >> 3) - 1.5f;
            float w = 0.f;
            if (gid == 0) w = 1.0f;
            else if (gid == 1) w = cx;
            else if (gid == 2) w = cy;
            else if (gid == 3) w = cx * cx;
            else if (gid == 4) w = cx * cy;
            else if (gid == 5) w = cy * cy;
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float t = __shfl_sync(FULL, vo[c], k);
                if (gid == c) v = t;
            }
            const uint32_t vh = to_tf32(v);
            const uint32_t vl = to_tf32(v - __uint_as_float(vh));
            if (h == 0) { aG0[s4] = to_tf32(w); aF1h[s4] = vh; aF1l[s4] = vl; }
            else { aG2[s4] = to_tf32(w); aF3h[s4] = vh; aF3l[s4] = vl; }
        }
    }

    const int wmax = __reduce_max_sync(FULL, last);
    if (lane == 0) s_wmax[warp] = wmax;
    __syncthreads();
    int max_last = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) max_last = max(max_last, s_wmax[w]);
    if (max_last == 0) return;

    float T = Tf;
    float D = 0.f;
    float* sbG = &s_b[warp][0][0][0];
    float* sbF = &s_b[warp][1][0][0];

    for (int hi = max_last; hi > 0; hi -= BLOCK_PIX) {
        const int lo = max(0, hi - BLOCK_PIX);
        const int cnt = hi - lo;
        __syncthreads();
        if (tid < cnt) {
            const int g = __ldg(ids + range.x + lo + tid);
            const float2 m = __ldg(reinterpret_cast<const float2*>(xy + int64_t(g) * st.xs));
            const float* cq = conic + int64_t(g) * st.cs;
            const float A = __ldg(cq), B = __ldg(cq + 1), Cc = __ldg(cq + 2);
            const float o = __ldg(opacity + int64_t(g) * st.os);
            float col[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c) col[c] = __ldg(colors + int64_t(g) * st.ks + c);
            s_rec[tid * 3 + 0] = make_float4(m.x, m.y, (-0.5f * LOG2E) * A, -LOG2E * B);
            s_rec[tid * 3 + 1] = make_float4((-0.5f * LOG2E) * Cc, o, col[0], col[1]);
            s_rec[tid * 3 + 2] = make_float4(col[2], col[3], __int_as_float(g), 0.f);
            s_mask[tid] = (unsigned char)block_mask(m.x, m.y, A, B, Cc, o, ox, oy);
        }
        __syncthreads();
        if (wmax <= lo) continue;
        const unsigned short* my_list = s_list[warp] + LIST_PAD;
        const int nl = build_list(s_mask, s_list[warp], min(cnt, wmax - lo), warp, lane);
        for (int ii = nl - 1; ii >= 0; ii -= GB) {
            unsigned present = 0;
#pragma unroll
            for (int u = 0; u < GB; ++u) {
                const int j = my_list[ii - u];
                const float4 r0 = s_rec[j * 3 + 0];
                const float4 r1 = s_rec[j * 3 + 1];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                const float p2 = fmaf(r1.x * dy, dy, fmaf(r0.w, dy, r0.z * dx) * dx);
                const float G = ex2_approx(p2);
                const float a = fminf(amax, r1.y * G);
                const bool valid = ((lo + j) < last) && !(p2 > 0.0f) && (a >= ALPHA_MIN);
                present |= (__ballot_sync(FULL, valid) != 0u) ? (1u << u) : 0u;
                float fac = 0.f, go = 0.f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - a);
                    T *= ra;
                    fac = a * T;
                    float S = r1.z * vo[0];
                    if (CH > 1) S = fmaf(r1.w, vo[1], S);
                    if (CH > 2) {
                        const float2 r2 = *reinterpret_cast<const float2*>(&s_rec[j * 3 + 2]);
                        S = fmaf(r2.x, vo[2], S);
                        if (CH > 3) S = fmaf(r2.y, vo[3], S);
                    }
                    const float v_al = fmaf(T, S, ra * (tail - D));
                    D = fmaf(fac, S, D);
                    if (!GSPLAT || (r1.y * G <= 0.999f)) go = G * v_al;
                }
                sbG[u * 32 + (lane ^ (4 * u))] = go;
                sbF[u * 32 + (lane ^ (4 * u))] = fac;
            }
            if (present == 0u) continue;
            __syncwarp();
            float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f}, acc3[4] = {0.f, 0.f, 0.f, 0.f};   // independent MMA chains
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int k0 = (8 * s4 + tig) ^ (4 * gid), k1 = (8 * s4 + tig + 4) ^ (4 * gid);
                {
                    const float b0 = sbG[gid * 32 + k0], b1 = sbG[gid * 32 + k1];
                    const uint32_t b0h = to_tf32(b0), b1h = to_tf32(b1);
                    const uint32_t b0l = to_tf32(b0 - __uint_as_float(b0h)), b1l = to_tf32(b1 - __uint_as_float(b1h));
                    mma_tf32(acc, aG0[s4], 0u, aG2[s4], 0u, b0h, b1h);
                    mma_tf32(acc2, aG0[s4], 0u, aG2[s4], 0u, b0l, b1l);
                }
                {
                    const float b0 = sbF[gid * 32 + k0], b1 = sbF[gid * 32 + k1];
                    const uint32_t b0h = to_tf32(b0), b1h = to_tf32(b1);
                    const uint32_t b0l = to_tf32(b0 - __uint_as_float(b0h)), b1l = to_tf32(b1 - __uint_as_float(b1h));
                    mma_tf32(acc3, 0u, aF1h[s4], 0u, aF3h[s4], b0h, b1h);
                    mma_tf32(acc2, 0u, aF1l[s4], 0u, aF3l[s4], b0h, b1h);
                    mma_tf32(acc, 0u, aF1h[s4], 0u, aF3h[s4], b0l, b1l);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += acc2[q] + acc3[q];
            __syncwarp();
            // acc[0], acc[1]: row gid (moment gid), entries 2*tig, 2*tig+1;  acc[2], acc[3]: row 8+gid (colour gid), same entries.
            // Collect the six moments and the colour sums of entries 2*tig+e in the gid==0 lane of each tig.
            float mom[2][6], csum[2][4];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                mom[0][r] = __shfl_sync(FULL, acc[0], (r << 2) | tig);
                mom[1][r] = __shfl_sync(FULL, acc[1], (r << 2) | tig);
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                csum[0][c] = __shfl_sync(FULL, acc[2], (c << 2) | tig);
                csum[1][c] = __shfl_sync(FULL, acc[3], (c << 2) | tig);
            }
            if (gid == 0) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = 2 * tig + e;
                    if (!((present >> n) & 1u)) continue;
                    const int j = my_list[ii - n];
                    const float4 r0 = s_rec[j * 3 + 0];
                    const float4 r1 = s_rec[j * 3 + 1];
                    const float A = r0.z * (-2.0f / LOG2E), B = r0.w * (-1.0f / LOG2E), Cc = r1.x * (-2.0f / LOG2E);
                    const int g = __float_as_int(s_rec[j * 3 + 2].z);
                    const float ex = r0.x - bcx, ey = r0.y - bcy;   // splat centre relative to the block centre; dx = ex - cx
                    const float S0 = mom[e][0], Sx = mom[e][1], Sy = mom[e][2], Sxx = mom[e][3], Sxy = mom[e][4], Syy = mom[e][5];
                    const float no = -r1.y;                          // v_sigma = -opacity * g
                    const float M1 = no * (ex * S0 - Sx), M2 = no * (ey * S0 - Sy);
                    const float M3 = no * (ex * ex * S0 - 2.0f * ex * Sx + Sxx);
                    const float M4 = no * (ex * ey * S0 - ex * Sy - ey * Sx + Sxy);
                    const float M5 = no * (ey * ey * S0 - 2.0f * ey * Sy + Syy);
                    float* vx = v_xy + int64_t(g) * st.xs;
                    float* vc = v_conic + int64_t(g) * st.cs;
                    atomicAdd(vx, (A * M1 + B * M2) * sx);
                    atomicAdd(vx + 1, (B * M1 + Cc * M2) * sy);
                    atomicAdd(vc, 0.5f * M3);
                    atomicAdd(vc + 1, M4);
                    atomicAdd(vc + 2, 0.5f * M5);
                    atomicAdd(v_opacity + int64_t(g) * st.os, S0);
#pragma unroll
                    for (int c = 0; c < CH; ++c) atomicAdd(v_colors + int64_t(g) * st.ks + c, csum[e][c]);
                }
            }
        }
    }
}

template <int CH>
int fwd_dispatch(int mode, int width, int height, const int32_t* ranges, const int32_t* ids, int row_stride, const float* xy, const float* conic,
                 const float* opacity, const float* colors, const float* bg, float* image, int64_t ps, int64_t cs, float* final_T,
                 int32_t* n_contrib, float* alpha, uint8_t* hit_any, cudaStream_t s) {
    const int gx = div_up(width, TILE), gy = div_up(height, TILE);
    dim3 grid(gx, gy);
    const SplatStrides st = row_stride > 0 ? SplatStrides{row_stride, row_stride, row_stride, row_stride} : SplatStrides{2, 3, 1, CH};
    // default: asynchronous (cp.async, double-buffered) slab staging; B200GS_FWD_SYNC=1 selects the synchronous kernel for A/B runs
    static const bool use_sync = []() { const char* e = getenv("B200GS_FWD_SYNC"); return e && e[0] == '1'; }();
    // the row layout is [x, y, depth, A, B, C, comp, opacity, r, g, b, radius] (include/b200gs.h); 16-byte copies need 16-byte aligned rows
    const bool rows16 = row_stride == 12 && CH == 3 && conic == xy + 3 && opacity == xy + 7 && colors == xy + 8 && (reinterpret_cast<uintptr_t>(xy) & 15) == 0;
    if (!use_sync || hit_any != nullptr) {
#define B200GS_FWD_ARGS width, height, gx, (const int2*)ranges, ids, st, xy, conic, opacity, colors, bg, image, ps, cs, final_T, n_contrib, alpha, hit_any
#define B200GS_FWD_LAUNCH(G, R)                                                                                                   \
    do {                                                                                                                           \
        if (hit_any) blend_fwd_async_kernel<CH, G, R, true><<<grid, BLOCK_PIX, 0, s>>>(B200GS_FWD_ARGS);                          \
        else blend_fwd_async_kernel<CH, G, R, false><<<grid, BLOCK_PIX, 0, s>>>(B200GS_FWD_ARGS);                                 \
    } while (0)
        if (rows16) {
            if constexpr (CH == 3) {
                if (mode == B200GS_MODE_GSPLAT) B200GS_FWD_LAUNCH(true, true);
                else B200GS_FWD_LAUNCH(false, true);
            }
        } else {
            if (mode == B200GS_MODE_GSPLAT) B200GS_FWD_LAUNCH(true, false);
            else B200GS_FWD_LAUNCH(false, false);
        }
#undef B200GS_FWD_LAUNCH
#undef B200GS_FWD_ARGS
        B200GS_LAUNCH_CHECK();
        return B200GS_OK;
    }
    if (mode == B200GS_MODE_GSPLAT)
        blend_fwd_kernel<CH, true><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, st, xy, conic,
                                                              opacity, colors, bg, image, ps, cs, final_T, n_contrib, alpha);
    else
        blend_fwd_kernel<CH, false><<<grid, BLOCK_PIX, 0, s>>>(width, height, gx, (const int2*)ranges, ids, st, xy, conic,
                                                               opacity, colors, bg, image, ps, cs, final_T, n_contrib, alpha);
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

#ifndef B200GS_BWD_RB
#define B200GS_BWD_RB 4
#endif

template <int CH>
int bwd_dispatch(int mode, int width, int height, const int32_t* ranges, const int32_t* ids, int row_stride, const float* xy, const float* conic,
                 const float* opacity, const float* colors, const float* bg, const float* final_T, const int32_t* n_contrib,
                 const float* v_image, int64_t ps, int64_t cs, const float* v_alpha, float sx, float sy, float* v_xy, float* v_conic,
                 float* v_opacity, float* v_colors, float* v_xy_abs, cudaStream_t s) {
    const int gx = div_up(width, TILE), gy = div_up(height, TILE);
    dim3 grid(gx, gy);
    constexpr int RB = B200GS_BWD_RB;
    const SplatStrides st = row_stride > 0 ? SplatStrides{row_stride, row_stride, row_stride, row_stride} : SplatStrides{2, 3, 1, CH};
#define B200GS_BWD_ARGS width, height, gx, (const int2*)ranges, ids, st, xy, conic, opacity, colors, bg, final_T, n_contrib, \
                        v_image, ps, cs, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs
#define B200GS_BWD_MMA_ARGS width, height, gx, (const int2*)ranges, ids, st, xy, conic, opacity, colors, bg, final_T, n_contrib, \
                            v_image, ps, cs, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors
    // opt-in (B200GS_BWD_MMA=1): measured 0.89 ms vs 0.69 ms for the shuffle butterfly at 1 M Gaussians / 1080p — the TF32 hi/lo
    // splits, the lane->fragment staging and the writer algebra cost as many issue slots as the butterfly saves (677 M vs
    // ~620 M warp instructions, ncu), so the shuffle kernel stays the default.
    static const bool use_mma = []() { const char* e = getenv("B200GS_BWD_MMA"); return e && e[0] == '1'; }();
    // opt-in (B200GS_BWD_VS=1): value-scatter reduction, see blend_bwd_kernel.  Written from the ncu instruction breakdown of
    // the default kernel at the end of round 1; NOT yet run on hardware — the default path does not use it.
    static const bool use_vs = []() { const char* e = getenv("B200GS_BWD_VS"); return e && e[0] == '1'; }();
    if (CH == 3 && RB == 4 && use_vs && !v_xy_abs) {
        if constexpr (CH == 3 && RB == 4) {
            if (mode == B200GS_MODE_GSPLAT)
                blend_bwd_kernel<CH, true, false, RB, true><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
            else
                blend_bwd_kernel<CH, false, false, RB, true><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
            B200GS_LAUNCH_CHECK();
            return B200GS_OK;
        }
    }
    // default: transpose-reduce kernel; B200GS_BWD_BUTTERFLY=1 selects the shuffle butterfly for A/B runs
    static const bool use_butterfly = []() { const char* e = getenv("B200GS_BWD_BUTTERFLY"); return e && e[0] == '1'; }();
    if (!use_butterfly && !use_mma) {
        static const cudaError_t attr_rc = []() {
            cudaError_t e = cudaSuccess;
            auto set = [&e](const void* f, size_t bytes) {
                if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            };
            set((const void*)blend_bwd_tr_kernel<CH, true, false>, sizeof(TrSmem<false>));
            set((const void*)blend_bwd_tr_kernel<CH, false, false>, sizeof(TrSmem<false>));
            set((const void*)blend_bwd_tr_kernel<CH, true, true>, sizeof(TrSmem<true>));
            set((const void*)blend_bwd_tr_kernel<CH, false, true>, sizeof(TrSmem<true>));
            return e;
        }();
        if (attr_rc != cudaSuccess) {
            set_error("blend_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_rc));
            return B200GS_ECUDA;
        }
        if (v_xy_abs) {
            if (mode == B200GS_MODE_GSPLAT)
                blend_bwd_tr_kernel<CH, true, true><<<grid, BLOCK_PIX, sizeof(TrSmem<true>), s>>>(B200GS_BWD_ARGS);
            else
                blend_bwd_tr_kernel<CH, false, true><<<grid, BLOCK_PIX, sizeof(TrSmem<true>), s>>>(B200GS_BWD_ARGS);
        } else {
            if (mode == B200GS_MODE_GSPLAT)
                blend_bwd_tr_kernel<CH, true, false><<<grid, BLOCK_PIX, sizeof(TrSmem<false>), s>>>(B200GS_BWD_ARGS);
            else
                blend_bwd_tr_kernel<CH, false, false><<<grid, BLOCK_PIX, sizeof(TrSmem<false>), s>>>(B200GS_BWD_ARGS);
        }
        B200GS_LAUNCH_CHECK();
        return B200GS_OK;
    }
    if (mode == B200GS_MODE_GSPLAT) {
        if (v_xy_abs)
            blend_bwd_kernel<CH, true, true, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
        else if (use_mma)
            blend_bwd_mma_kernel<CH, true><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_MMA_ARGS);
        else
            blend_bwd_kernel<CH, true, false, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
    } else {
        if (v_xy_abs)
            blend_bwd_kernel<CH, false, true, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
        else if (use_mma)
            blend_bwd_mma_kernel<CH, false><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_MMA_ARGS);
        else
            blend_bwd_kernel<CH, false, false, RB><<<grid, BLOCK_PIX, 0, s>>>(B200GS_BWD_ARGS);
    }
#undef B200GS_BWD_MMA_ARGS
#undef B200GS_BWD_ARGS
    B200GS_LAUNCH_CHECK();
    return B200GS_OK;
}

}  // namespace

int launch_blend_fwd(int mode, int width, int height, int channels, const int32_t* ranges, const int32_t* ids, int row_stride, const float* xy,
                     const float* conic, const float* opacity, const float* colors, const float* bg, float* image,
                     int64_t pix_stride, int64_t ch_stride, float* final_T, int32_t* n_contrib, float* alpha, cudaStream_t s,
                     uint8_t* hit_any) {
    switch (channels) {
        case 1: return fwd_dispatch<1>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, hit_any, s);
        case 2: return fwd_dispatch<2>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, hit_any, s);
        case 3: return fwd_dispatch<3>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, hit_any, s);
        case 4: return fwd_dispatch<4>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, image, pix_stride, ch_stride, final_T, n_contrib, alpha, hit_any, s);
    }
    set_error("blend_fwd: unsupported channel count %d (1..4)", channels);
    return B200GS_EINVAL;
}

int launch_blend_bwd(int mode, int width, int height, int channels, const int32_t* ranges, const int32_t* ids, int row_stride, const float* xy,
                     const float* conic, const float* opacity, const float* colors, const float* bg, const float* final_T,
                     const int32_t* n_contrib, const float* v_image, int64_t pix_stride, int64_t ch_stride, const float* v_alpha,
                     float sx, float sy, float* v_xy, float* v_conic, float* v_opacity, float* v_colors, float* v_xy_abs,
                     cudaStream_t s) {
    switch (channels) {
        case 1: return bwd_dispatch<1>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 2: return bwd_dispatch<2>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 3: return bwd_dispatch<3>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
        case 4: return bwd_dispatch<4>(mode, width, height, ranges, ids, row_stride, xy, conic, opacity, colors, bg, final_T, n_contrib, v_image, pix_stride, ch_stride, v_alpha, sx, sy, v_xy, v_conic, v_opacity, v_colors, v_xy_abs, s);
    }
    set_error("blend_bwd: unsupported channel count %d (1..4)", channels);
    return B200GS_EINVAL;
}

}  // namespace b200gs
