// microbench.cu -- measured integer-pipe peaks: the roofline denominators.
//
// the DP relaxation is bound by the 32-bit integer / DPX issue
// rate, and MEASURED_PEAKS.json carries no integer figure, so the build
// measures one on the GPU it runs on: dependency-free (8 independent chains per
// thread, 32 resident warps per SM) loops of
//   kind 0  VIADDMNMX + VIMNMX3 only              (pure DPX, ALU pipe)
//   kind 1  the 5-op linear-gap cell mix           ISETP SEL IADD VIMNMX VIADDMNMX
//   kind 2  the 7-op affine-gap cell mix           ISETP SEL IADD 2xVIADDMNMX VIMNMX3 IADD
//   kind 3  IMAD only                              (FMA pipe)
//   kind 4  kind 0 and kind 3 interleaved 1:1      (do the pipes dual-issue?)
//   kind 5  the strip kernel's own Gotoh cell      2xVIADDMNMX VIMNMX3 + R2P/7 | 3x IMAD
//   kind 6  the strip kernel's own linear cell     VIMNMX VIADDMNMX + R2P/7    | 2x IMAD
//           (kinds 5/6 return CELLS per second: the cell-update peak of the mix
//            the kernel really executes, both pipes busy)
// The result is lane-operations per second over the whole chip.
#include "engine.cuh"
#include "strip_kernel.cuh"

namespace anyseq {

constexpr int kChains = 8;

// kinds 5 / 6: 32 cells of the strip kernel's own instruction sequence (8
// independent chains x 4), the match predicate taken from bit IDX of a mask
// exactly as in strip_kernel.cuh (ptxas folds the bit tests into R2P)
template <int KIND, int IDX>
struct MixCell {
    static __device__ __forceinline__ void run(int (&a)[kChains], int (&b)[kChains], int (&c)[kChains],
                                               int (&d)[kChains], unsigned mask, int p0, int p1, int p2, int p3)
    {
        constexpr int k = IDX % kChains;
        if constexpr (KIND == 5) {
            const int dd = diag_plus_sigma_mask<IDX>(mask, b[k], p3, p0, p1);
            const int e = __viaddmax_s32(a[k], p2, b[k]);
            const int f = __viaddmax_s32(d[k], p2, c[k]);
            const int h = __vimax3_s32(dd, e, f);
            a[k] = e; d[k] = f; c[k] = b[k]; b[k] = imad_add(h, p3, p2);
        } else {
            const int dd = diag_plus_sigma_mask<IDX>(mask, a[k], p3, p0, p1);
            const int t = max(b[k], d[k]);
            const int h = __viaddmax_s32(t, p2, dd);
            c[k] = a[k]; a[k] = d[k]; d[k] = b[k]; b[k] = h;
        }
        if constexpr (IDX + 1 < 4 * kChains) MixCell<KIND, IDX + 1>::run(a, b, c, d, mask, p0, p1, p2, p3);
    }
};

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                       int iters, int p0, int p1, int p2, int p3,
                                                       long long* clocks)
{
    int a[kChains], b[kChains], c[kChains], d[kChains];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
        a[k] = in[(tid + k) & 1023];
        b[k] = in[(tid + 2 * k + 1) & 1023];
        c[k] = in[(tid + 3 * k + 2) & 1023];
        d[k] = in[(tid + 5 * k + 3) & 1023];
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if constexpr (KIND == 5 || KIND == 6) {
            MixCell<KIND, 0>::run(a, b, c, d, (unsigned)c[0] ^ (unsigned)it, p0, p1, p2, p3);
        } else
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) {
                if constexpr (KIND == 0) {
                    a[k] = __viaddmax_s32(a[k], p0, b[k]);
                    b[k] = __vimax3_s32(b[k], c[k], a[k]);
                } else if constexpr (KIND == 1) {
                    // linear cell: h = max(max(left, up) + g, diag + sigma)
                    const int sub = (c[k] == d[k]) ? p0 : p1;
                    const int dd = a[k] + sub;
                    const int t = max(b[k], a[k]);
                    const int h = __viaddmax_s32(t, p2, dd);
                    a[k] = b[k]; b[k] = h; d[k] = d[k] ^ (h & 3);
                } else if constexpr (KIND == 2) {
                    // affine cell in X form (strip_kernel.cuh)
                    const int sub = (c[k] == (d[k] & 3)) ? p0 : p1;
                    const int e = __viaddmax_s32(a[k], p2, b[k]);
                    const int f = __viaddmax_s32(d[k], p2, c[k]);
                    const int dd = b[k] + sub;
                    const int h = __vimax3_s32(dd, e, f);
                    a[k] = e; d[k] = f; b[k] = h + p3;
                } else if constexpr (KIND == 3) {
                    a[k] = a[k] * p0 + b[k];
                    b[k] = b[k] * p1 + a[k];
                } else {
                    a[k] = __viaddmax_s32(a[k], p0, c[k]);
                    b[k] = b[k] * p1 + d[k];
                }
            }
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc += a[k] ^ b[k] ^ c[k] ^ d[k];
    out[tid] = acc;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

static int ops_per_inner(int kind)
{
    switch (kind) {
        case 0: return 2;
        case 1: return 5;   // counted as the algorithmic 5 (the xor feeding d is bookkeeping)
        case 2: return 7;
        case 3: return 2;
        case 5: return 1;   // cells
        case 6: return 1;   // cells
        default: return 2;
    }
}

int Engine::measure_int_peak(int kind, double* ops_per_s, float* sm_mhz)
{
    std::lock_guard<std::recursive_mutex> lock(mu_);
    ANYSEQ_CUDA_CHECK(cudaSetDevice(device));
    if (kind < 0 || kind > 6) return ANYSEQ_ERR_BAD_ARG;
    const int blocks = sm_count * 8, threads = 256;
    const int iters = 20000;
    if (aux_.ensure(sizeof(int) * 1024 + sizeof(int) * (size_t)blocks * threads + sizeof(long long) * blocks + 64))
        return ANYSEQ_ERR_NO_DEVICE;
    int* d_in = aux_.as<int>();
    int* d_out = d_in + 1024;
    long long* d_clk = reinterpret_cast<long long*>(d_out + (size_t)blocks * threads + ((size_t)blocks * threads % 2));
    std::vector<int> h_in(1024);
    for (int i = 0; i < 1024; ++i) h_in[i] = (i * 2654435761u) >> 20;
    ANYSEQ_CUDA_CHECK(cudaMemcpyAsync(d_in, h_in.data(), sizeof(int) * 1024, cudaMemcpyHostToDevice, stream_));
    auto launch = [&](int n) {
        switch (kind) {
            case 0: int_peak_kernel<0><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, -1, 2, -3, -2, d_clk); break;
            case 1: int_peak_kernel<1><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, 2, -1, -1, 0, d_clk); break;
            case 2: int_peak_kernel<2><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, 2, -1, -1, -3, d_clk); break;
            case 3: int_peak_kernel<3><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, 3, 5, 0, 0, d_clk); break;
            case 5: int_peak_kernel<5><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, -1, 2, -1, 1, d_clk); break;
            case 6: int_peak_kernel<6><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, -1, 2, -1, 1, d_clk); break;
            default: int_peak_kernel<4><<<blocks, threads, 0, stream_>>>(d_in, d_out, n, -1, 3, 0, 0, d_clk); break;
        }
    };
    launch(iters / 10);   // warm-up
    ANYSEQ_CUDA_CHECK(cudaGetLastError());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        ANYSEQ_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
        launch(iters);
        ANYSEQ_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
        ANYSEQ_CUDA_CHECK(cudaStreamSynchronize(stream_));
        float ms = 0.f;
        ANYSEQ_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
        best_ms = std::min(best_ms, ms);
    }
    long long clk = 0;
    ANYSEQ_CUDA_CHECK(cudaMemcpy(&clk, d_clk, sizeof(long long), cudaMemcpyDeviceToHost));
    const double lane_ops = (double)blocks * threads * (double)iters * 4.0 * kChains * ops_per_inner(kind);
    if (ops_per_s) *ops_per_s = lane_ops / (best_ms * 1e-3);
    if (sm_mhz) *sm_mhz = (float)((double)clk / (best_ms * 1e-3) / 1e6);
    return ANYSEQ_OK;
}

}  // namespace anyseq
