// Micro-benchmark of the hand-written onesweep digit pass (csrc/onesweep.cuh): times one pass over n random records for a few
// look-back configurations and prints, from a traced launch, where a tile's time goes (globaltimer at the phase boundaries).
// Build + run on the GPU box:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I gaussian-splatting-lightning_b200/csrc
//                              -I include profiles/tools/sweep_bench.cu -o /tmp/sweep_bench && /tmp/sweep_bench
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "onesweep.cuh"

namespace b200gs {
thread_local char g_err[8] = {0};
void set_error(const char*, ...) {}
void count_launch() {}
}  // namespace b200gs

using namespace b200gs;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <typename Rec> __host__ __device__ uint32_t& key_ref(Rec& r);
template <> __host__ __device__ uint32_t& key_ref<uint2>(uint2& r) { return r.x; }
template <> __host__ __device__ uint32_t& key_ref<uint4>(uint4& r) { return r.w; }

template <typename Rec, int IPT, int LB0, int LBMAX, int LDMODE>
void run(const char* label, int64_t n, int digit_bits, int reps) {
    constexpr int TILE = sweep::PASS_THREADS * IPT;
    const int64_t tiles = (n + TILE - 1) / TILE;
    std::vector<Rec> h(n);
    srand(1);
    for (int64_t i = 0; i < n; ++i) {
        Rec r{};
        key_ref(r) = ((uint32_t)rand() * 2654435761u) >> (32 - digit_bits);
        h[i] = r;
    }
    std::vector<uint32_t> hist(256, 0);
    for (int64_t i = 0; i < n; ++i) hist[key_ref(h[i]) & 255u]++;
    Rec *in, *out;
    uint32_t *d_hist, *lookback, *ticket;
    unsigned long long* trace;
    CK(cudaMalloc(&in, n * sizeof(Rec)));
    CK(cudaMalloc(&out, n * sizeof(Rec)));
    CK(cudaMalloc(&d_hist, 256 * 4));
    CK(cudaMalloc(&lookback, tiles * 256 * 4));
    CK(cudaMalloc(&ticket, 64));
    CK(cudaMalloc(&trace, tiles * 8 * 8));
    CK(cudaMemcpy(in, h.data(), n * sizeof(Rec), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_hist, hist.data(), 256 * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> ms;
    for (int r = 0; r < reps + 3; ++r) {
        CK(cudaMemsetAsync(lookback, 0, tiles * 256 * 4));
        CK(cudaMemsetAsync(ticket, 0, 64));
        CK(cudaEventRecord(e0));
        sweep::onesweep_pass_kernel<Rec, IPT, LB0, LBMAX, LDMODE, false><<<(unsigned)tiles, sweep::PASS_THREADS>>>(in, out, nullptr, n, 0, d_hist, lookback, ticket);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        if (r >= 3) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    // correctness: stable partition by the low byte
    std::vector<Rec> got(n);
    CK(cudaMemcpy(got.data(), out, n * sizeof(Rec), cudaMemcpyDeviceToHost));
    std::vector<Rec> ref(n);
    {
        std::vector<int64_t> off(257, 0);
        for (int d = 0; d < 256; ++d) off[d + 1] = off[d] + hist[d];
        for (int64_t i = 0; i < n; ++i) ref[off[key_ref(h[i]) & 255u]++] = h[i];
    }
    bool ok = true;
    for (int64_t i = 0; i < n && ok; ++i) ok = key_ref(got[i]) == key_ref(ref[i]);
    // traced launch
    CK(cudaMemset(lookback, 0, tiles * 256 * 4));
    CK(cudaMemset(ticket, 0, 64));
    sweep::onesweep_pass_kernel<Rec, IPT, LB0, LBMAX, LDMODE, true><<<(unsigned)tiles, sweep::PASS_THREADS>>>(in, out, nullptr, n, 0, d_hist, lookback, ticket, trace);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> tr(tiles * 8);
    CK(cudaMemcpy(tr.data(), trace, tiles * 64, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int64_t t = 0; t < tiles; ++t) t0 = std::min(t0, tr[t * 8]);
    double mx[6] = {0}, av[6] = {0};
    for (int64_t t = 0; t < tiles; ++t)
        for (int k = 0; k < 6; ++k) {
            const double v = (double)(tr[t * 8 + k] - t0) * 1e-3;
            mx[k] = std::max(mx[k], v);
            av[k] += v / tiles;
        }
    printf("%-34s n=%8lld tiles=%5lld digits=%3d  pass median %.2f us  min %.2f  %s\n", label, (long long)n, (long long)tiles, 1 << digit_bits,
           ms[ms.size() / 2] * 1e3, ms[0] * 1e3, ok ? "ok" : "WRONG");
    printf("    phase ends (us since first tile start), avg / max over tiles: start %.1f/%.1f  ticket %.1f/%.1f  ranked %.1f/%.1f  published+staged %.1f/%.1f  "
           "looked-back %.1f/%.1f  written %.1f/%.1f\n", av[0], mx[0], av[1], mx[1], av[2], mx[2], av[3], mx[3], av[4], mx[4], av[5], mx[5]);
    // the last tiles' own timelines
    for (int64_t t : {tiles / 4, tiles / 2, tiles - 1}) {
        printf("    tile %5lld:", (long long)t);
        for (int k = 0; k < 6; ++k) printf(" %.1f", (double)(tr[t * 8 + k] - t0) * 1e-3);
        printf("\n");
    }
    cudaFree(in); cudaFree(out); cudaFree(d_hist); cudaFree(lookback); cudaFree(ticket); cudaFree(trace);
}

int main() {
    const int64_t sizes[2] = {646000, 2300000};
    for (int64_t n : sizes) {
        run<uint2, 8, 8, 8, 0>("uint2 fixed 8 volatile", n, 8, 30);
        run<uint2, 8, 4, 32, 0>("uint2 4..32 volatile", n, 8, 30);
        run<uint2, 8, 8, 8, 1>("uint2 fixed 8 relaxed.gpu pred", n, 8, 30);
        run<uint2, 8, 4, 32, 1>("uint2 4..32 relaxed.gpu pred", n, 8, 30);
        run<uint2, 8, 16, 32, 1>("uint2 16..32 relaxed.gpu pred", n, 8, 30);
        run<uint2, 8, 32, 32, 1>("uint2 fixed 32 relaxed.gpu pred", n, 8, 30);
        run<uint2, 4, 4, 32, 1>("uint2 IPT4 4..32 relaxed pred", n, 8, 30);
        run<uint2, 8, 4, 32, 1>("uint2 4..32 relaxed, 4 digits", n, 2, 30);
    }
    run<uint4, 4, 8, 8, 0>("uint4 fixed 8 volatile", 1330000, 8, 30);
    run<uint4, 4, 4, 32, 1>("uint4 4..32 relaxed pred", 1330000, 8, 30);
    run<uint4, 4, 4, 32, 1>("uint4 4..32 relaxed pred (135 cells)", 1330000, 7, 30);
    run<uint4, 4, 4, 32, 1>("uint4 4..32 relaxed pred 4.4M", 4400000, 8, 30);
    return 0;
}
