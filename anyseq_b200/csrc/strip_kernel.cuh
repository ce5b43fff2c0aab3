// strip_kernel.cuh -- the DP relaxation hot path, hand-written for sm_100a.
//
// Replaces the reference's iteration_{cpu,acc} / scoring_{cpu,acc} /
// mapping_{cpu,acc} layer (src/iteration_acc.impala:16-172,
// src/scoring_acc.impala:1-181, src/mapping_acc.impala) -- same recurrences
// (relax_global / relax_local, src/align.impala:46-79), different machine
// mapping:
//
//   * one WARP owns a column strip of 32*K subject columns; each lane keeps K
//     columns of H (and F for Gotoh) in registers for the whole band,
//   * lane l works on row t-l at step t (anti-diagonal across lanes); the
//     right edge (H, E) moves to lane l+1 with warp shuffles,
//   * the cell update is DPX: VIADDMNMX / VIMNMX3(.RELU) (one instruction per
//     max(a+b,c) / max(a,b,c)),
//   * strips are chained through HBM/L2: lane 31's edge is staged in shared
//     memory and published 32 rows at a time (coalesced, .cg) with a
//     release/acquire row counter per strip -- no kernel relaunch, no grid
//     barrier (the reference relaunches per block anti-diagonal),
//   * a persistent grid takes (band, strip) items in dependency order, so all
//     waits are on lower-numbered items: deadlock-free when co-resident.
//
// Score domain: signed 32-bit, identical to the reference (Score = i32,
// src/dynprog.impala:10).  For Gotoh the registers hold X = H + gap_open so
// that E and F need one VIADDMNMX each:  E' = max(E + ge, X_left).
#pragma once

#include <type_traits>

#include "common.cuh"

namespace anyseq {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * kWarp;
constexpr unsigned kFull = 0xffffffffu;

struct KernelArgs {
    const Job* jobs;
    int njobs;
    long long total_items;
    ScoreParams sp;
    int one;                        // == 1, opaque to the compiler (see imad_add)
    int* status;                    // [0] StatusCode, [1..3] diagnostics
    unsigned long long timeout_ns;  // watchdog for dependency waits
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(int* p, int v)
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// a + b on the FMA pipe: IMAD with a multiplier the compiler cannot fold (one == 1
// at run time).  The B200 issues IMAD (FMA pipe) and DPX/compare ops (ALU pipe)
// side by side at 64 lanes/clk/SM each (measured: microbench.cu kind 4), so every
// add moved here is an ALU slot freed for VIADDMNMX / VIMNMX3.
__device__ __forceinline__ int imad_add(int a, int one, int b)
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
// d + (q == s ? same : diff): one compare (ALU) and two IMADs, the second one
// predicated -- no SEL.
__device__ __forceinline__ int diag_plus_sigma(int qc, int sc, int d, int one, int diff, int same)
{
    int dd;
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tmad.lo.s32 %0, %3, %4, %5;\n\t@p mad.lo.s32 %0, %3, %4, %6;\n\t}"
        : "=&r"(dd)
        : "r"(qc), "r"(sc), "r"(d), "r"(one), "r"(diff), "r"(same));
    return dd;
}

// Wait until *flag >= need.  Executed by all lanes of a warp (same address:
// one transaction).  Returns false if the watchdog fired or another warp
// already reported a failure -- the caller then leaves the kernel, so a logic
// error shows up as a status code instead of a hung GPU.
__device__ __forceinline__ bool wait_rows(const int* flag, int need, bool sys, int* status,
                                          unsigned long long timeout_ns)
{
    bool ok = true;
    int v = sys ? ld_acquire_sys(flag) : ld_acquire_gpu(flag);
    if (v < need) {
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (true) {
            v = sys ? ld_acquire_sys(flag) : ld_acquire_gpu(flag);
            if (v >= need) break;
            __nanosleep(100);
            if ((++spins & 127u) == 0u) {
                if (*(volatile int*)status != kStatusOk) { ok = false; break; }
                if (global_timer_ns() - t0 > timeout_ns) {
                    if (atomicCAS(status, kStatusOk, kStatusTimeout) == kStatusOk) {
                        status[1] = need; status[2] = v;
                    }
                    ok = false;
                    break;
                }
            }
        }
    }
    return __all_sync(kFull, ok);
}

template <int K>
__device__ __forceinline__ void load_row_ints(const int* __restrict__ p, int (&dst)[K])
{
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int c = 0; c < K; c += 4) {
            const int4 v = __ldcg(reinterpret_cast<const int4*>(p + c));
            dst[c] = v.x; dst[c + 1] = v.y; dst[c + 2] = v.z; dst[c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < K; ++c) dst[c] = __ldcg(p + c);
    }
}

// One (band, strip) item.  PARTIAL = the strip is cut by the right matrix edge
// (only the last strip of a job can be): the columns past the edge compute
// don't-care values (dependencies only run left->right, so they never reach a
// valid cell), the edge column is picked out for the output, and the local
// maximum is masked.
//
// Instruction budget per cell (SASS, checked with cuobjdump):
//   Gotoh : ISETP, VIADDMNMX (E), VIADDMNMX (F), VIMNMX3[.RELU]   -> 4 ALU
//           IMAD, @p IMAD (diag + sigma), IMAD (X = H + open)     -> 3 FMA
//   linear: ISETP, VIMNMX, VIADDMNMX[.RELU]                       -> 3 ALU
//           IMAD, @p IMAD                                         -> 2 FMA
template <bool LOCAL, bool AFFINE, int K, bool PARTIAL>
__device__ __forceinline__ bool process_item(const Job& J, const int band, const int strip,
                                             const ScoreParams& sp, const int one,
                                             int2* __restrict__ s_in, int2* __restrict__ s_out,
                                             uint8_t* __restrict__ s_q, const int lane, int* status,
                                             const unsigned long long timeout_ns)
{
    constexpr int SW = kWarp * K;
    const int i0 = band * J.band_h;
    const int hb = min(J.band_h, J.h - i0);
    const int j0 = strip * SW;
    const int wv = min(SW, J.w - j0);
    const bool last_strip = (strip == J.nstrips - 1);
    const int go = AFFINE ? sp.gap_open : 0;
    const int ge = sp.gap_extend;
    const int same_o = sp.same - go, diff_o = sp.diff - go;

    // left border source (rows of this band)
    const int* linH;
    const int* linE;
    const int* lflag;
    bool lsys = false;
    if (strip == 0) {
        linH = (J.inH ? J.inH : J.colH) + i0;
        linE = AFFINE ? (J.inH ? J.inE : J.colE) + i0 : nullptr;
        lflag = J.in_progress;
        lsys = true;
    } else {
        linH = J.colH + i0;
        linE = AFFINE ? J.colE + i0 : nullptr;
        lflag = J.progress + (strip - 1);
    }
    const bool mirror = last_strip && (J.outH != nullptr);
    const uint8_t* qrow = J.q + i0;

    // the band above must be complete (its bottom border is our top border)
    if (band > 0) {
        if (!wait_rows(J.progress + strip, i0, false, status, timeout_ns)) return false;
    }

    int X[K], F[K], sc[K];
    const int jl = j0 + lane * K;
    load_row_ints<K>(J.rowH + jl, X);
    if constexpr (AFFINE) {
        load_row_ints<K>(J.rowF + jl, F);
#pragma unroll
        for (int c = 0; c < K; ++c) X[c] += go;
    }
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int j = jl + c;
        sc[c] = (j < J.w) ? (int)J.s[j] : 0x7fff;   // never equals a byte
    }

    // H(i0-1, first column - 1) in X form
    int dcarry = __shfl_up_sync(kFull, X[K - 1], 1);
    if (lane == 0) dcarry = __ldcg(J.corner + strip) + go;

    const int outlane = PARTIAL ? (wv - 1) / K : 31;
    const int outc = PARTIAL ? (wv - 1) % K : K - 1;
    const int nvalid = PARTIAL ? max(0, min(K, wv - lane * K)) : K;
    const int T = hb + outlane;            // number of steps
    int hr = 0, er = 0;
    int best = kScoreMin;
    int flushed = 0;

    auto flush32 = [&](int base) {
        // rows [base, base+32) of the out lane's edge -> colH/colE (coalesced)
        const int r = base + lane;
        if (r < hb) {
            const int2 v = s_out[r & 63];
            __stcg(J.colH + i0 + r, v.x - go);
            if constexpr (AFFINE) __stcg(J.colE + i0 + r, v.y);
            if (mirror) {
                J.outH[i0 + r] = v.x - go;
                if constexpr (AFFINE) J.outE[i0 + r] = v.y;
            }
        }
    };
    auto publish = [&](int rows_abs) {
        __syncwarp();
        if (lane == 0) {
            if (mirror) {
                __threadfence_system();
                st_relaxed_sys(J.out_progress, rows_abs);
            }
            __threadfence();
            st_relaxed_gpu(J.progress + strip, rows_abs);
        }
    };

    // one anti-diagonal step; GUARD = some lanes may be outside [0, hb)
    auto step = [&](auto guard_tag, const int t) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        int xl = __shfl_up_sync(kFull, hr, 1);
        int el = 0;
        if constexpr (AFFINE) el = __shfl_up_sync(kFull, er, 1);
        const int2 bnd = s_in[t & 31];
        if (lane == 0) { xl = bnd.x; el = bnd.y; }
        const int i = t - lane;
        if (!GUARD || (unsigned)i < (unsigned)hb) {
            const int qc = s_q[i & 63];
            int d = dcarry;
            dcarry = xl;
            int xleft = xl;
            int e = el;
            int es = 0;   // PARTIAL: E of the edge column
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const int up = X[c];
                int h;
                if constexpr (AFFINE) {
                    const int dd = diag_plus_sigma(qc, sc[c], d, one, diff_o, same_o);
                    e = __viaddmax_s32(e, ge, xleft);
                    if constexpr (PARTIAL) {
                        if (c == outc) es = e;
                    }
                    const int f = __viaddmax_s32(F[c], ge, up);
                    h = LOCAL ? __vimax3_s32_relu(dd, e, f) : __vimax3_s32(dd, e, f);
                    F[c] = f;
                } else {
                    const int dd = diag_plus_sigma(qc, sc[c], d, one, sp.diff, sp.same);
                    const int tmax = max(xleft, up);
                    h = LOCAL ? __viaddmax_s32_relu(tmax, ge, dd) : __viaddmax_s32(tmax, ge, dd);
                }
                if constexpr (LOCAL) {
                    if (!PARTIAL || c < nvalid) best = max(best, h);
                }
                const int x = AFFINE ? imad_add(h, one, go) : h;
                d = up;
                X[c] = x;
                xleft = x;
            }
            hr = xleft;
            er = e;
            if constexpr (PARTIAL) {
                if (lane == outlane) {
                    int hs = X[0];
#pragma unroll
                    for (int c = 1; c < K; ++c)
                        if (c == outc) hs = X[c];
                    s_out[i & 63] = make_int2(hs, es);
                }
            } else {
                if (lane == 31) s_out[i & 63] = make_int2(hr, er);
            }
        }
    };

    for (int tb = 0; tb < T; tb += 32) {
        __syncwarp();
        // (1) publish the edge rows the out lane finished so far
        {
            const int completed = min(max(tb - outlane, 0), hb);
            bool any = false;
            while (completed - flushed >= 32) {
                flush32(flushed);
                flushed += 32;
                any = true;
            }
            if (any && flushed < hb) publish(i0 + flushed);   // the final publish is below
        }
        // (2) fetch the next 32 rows of the left border and of the query
        {
            const int r = tb + lane;
            int2 v = make_int2(0, kNegInf);
            uint8_t qv = 0;
            if (tb < hb) {
                if (lflag != nullptr) {
                    if (!wait_rows(lflag, i0 + min(tb + 32, hb), lsys, status, timeout_ns)) return false;
                }
                if (r < hb) {
                    v.x = __ldcg(linH + r) + go;
                    if constexpr (AFFINE) v.y = __ldcg(linE + r);
                    qv = qrow[r];
                }
            }
            s_in[lane] = v;
            s_q[r & 63] = qv;
            __syncwarp();
        }
        // (3) 32 anti-diagonal steps; batches in which every lane is inside the
        //     band run the unguarded body
        if (tb >= 31 && tb + 32 <= hb) {
#pragma unroll 1
            for (int t = tb; t < tb + 32; ++t) step(std::false_type{}, t);
        } else {
            const int tend = min(tb + 32, T);
#pragma unroll 1
            for (int t = tb; t < tend; ++t) step(std::true_type{}, t);
        }
    }

    // drain: remaining edge rows, bottom border, corner, local maximum
    __syncwarp();
    while (flushed < hb) {
        flush32(flushed);
        flushed += 32;
    }
    {
        if constexpr (K % 4 == 0) {
#pragma unroll
            for (int c = 0; c < K; c += 4) {
                __stcg(reinterpret_cast<int4*>(J.rowH + jl + c),
                       make_int4(X[c] - go, X[c + 1] - go, X[c + 2] - go, X[c + 3] - go));
                if constexpr (AFFINE)
                    __stcg(reinterpret_cast<int4*>(J.rowF + jl + c),
                           make_int4(F[c], F[c + 1], F[c + 2], F[c + 3]));
            }
        } else {
#pragma unroll
            for (int c = 0; c < K; ++c) {
                __stcg(J.rowH + jl + c, X[c] - go);
                if constexpr (AFFINE) __stcg(J.rowF + jl + c, F[c]);
            }
        }
        if (lane == 0) __stcg(J.corner + strip, dcarry - go);
    }
    if constexpr (LOCAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(kFull, best, o));
        if (lane == 0) atomicMax(J.best, best);
    }
    publish(i0 + hb);
    return true;
}

// registers per thread are capped so that 4 CTAs (16 warps, 4 per scheduler)
// fit on an SM for every K: the recurrence has a 3-deep dependent chain per cell
// and needs that many warps to keep the ALU pipe busy.
template <bool LOCAL, bool AFFINE, int K>
__global__ void __launch_bounds__(kThreads, (K >= 32 ? 4 : (K >= 16 ? 5 : 6)))
strip_kernel(const KernelArgs a)
{
    __shared__ int2 s_in[kWarpsPerBlock][32];
    __shared__ int2 s_out[kWarpsPerBlock][64];
    __shared__ uint8_t s_q[kWarpsPerBlock][64];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    constexpr int SW = kWarp * K;

    int jcur = 0;
    for (long long item = (long long)blockIdx.x * kWarpsPerBlock + warp; item < a.total_items;
         item += nwarps) {
        while (jcur + 1 < a.njobs && item >= a.jobs[jcur + 1].item_begin) ++jcur;
        const Job& J = a.jobs[jcur];
        const long long loc = item - J.item_begin;
        const int band = (int)(loc / J.nstrips);
        const int strip = (int)(loc % J.nstrips);
        const bool partial = (strip + 1) * SW > J.w;
        bool ok;
        if (partial)
            ok = process_item<LOCAL, AFFINE, K, true>(J, band, strip, a.sp, a.one, s_in[warp],
                                                      s_out[warp], s_q[warp], lane, a.status,
                                                      a.timeout_ns);
        else
            ok = process_item<LOCAL, AFFINE, K, false>(J, band, strip, a.sp, a.one, s_in[warp],
                                                       s_out[warp], s_q[warp], lane, a.status,
                                                       a.timeout_ns);
        if (!ok) return;
    }
}

}  // namespace anyseq
