// batch.cuh -- argument block shared by the batch kernels (batch.cu, batch_x2.cu)
#pragma once
#include "common.cuh"

namespace anyseq {

struct BatchArgs {
    const uint8_t* q;
    const long long* qoff;
    const uint8_t* s;
    const long long* soff;
    long long npairs;
    int* scores;
    ScoreParams sp;
    int gap_init;
    int mode;
    int one;
    int ncodes;
    int cols_longer;               // 1: columns = longer sequence of a pair, 0: shorter
    const uint8_t* lut;            // byte -> code (shared by both sequences)
    unsigned long long* counter;   // next pair (packed kernels: pair of pairs) to claim
    int bias;                      // packed 16-bit kernels: offset that keeps every stored value non-negative
    // 2-bit packed DNA input (anyseq_score_batch_packed2): four symbols per byte, least significant bits first, every
    // sequence starting on a byte boundary.  q / s then hold packed bytes, qoff / soff BYTE offsets of the sequences
    // (nullptr: sequence p starts at p * stride) and the lengths come from qlen / slen (nullptr: uniform length).
    int packed2;
    const int* qlen;
    const int* slen;
    int qlen_u, slen_u;
    long long qstride, sstride;
};

// where pair p's sequences start (symbol offset, or byte offset for 2-bit input) and how many symbols they have
__device__ __forceinline__ long long batch_q_start(const BatchArgs& a, long long p) { return a.qoff ? a.qoff[p] : p * a.qstride; }
__device__ __forceinline__ long long batch_s_start(const BatchArgs& a, long long p) { return a.soff ? a.soff[p] : p * a.sstride; }
__device__ __forceinline__ int batch_q_len(const BatchArgs& a, long long p)
{
    return a.packed2 ? (a.qlen ? a.qlen[p] : a.qlen_u) : (int)(a.qoff[p + 1] - a.qoff[p]);
}
__device__ __forceinline__ int batch_s_len(const BatchArgs& a, long long p)
{
    return a.packed2 ? (a.slen ? a.slen[p] : a.slen_u) : (int)(a.soff[p + 1] - a.soff[p]);
}
// alphabet code of symbol idx of a sequence: byte -> code table, or the 2-bit value + 1 (code 0 matches nothing)
__device__ __forceinline__ int batch_code(const uint8_t* __restrict__ base, int idx, int packed2, const uint8_t* lut)
{
    if (packed2) return (int)((base[idx >> 2] >> (2 * (idx & 3))) & 3u) + 1;
    return (int)lut[base[idx]];
}

using BatchKernelFn = void (*)(const BatchArgs);
// packed 16-bit kernels (batch_x2.cu): two pairs of equal shape per warp; nullptr if K is not instantiated
BatchKernelFn pick_batch_x2_kernel(int mode, bool affine, int K);
// four pairs per warp (16 lanes x 32 columns per pair of pairs): column sequences of at most 512 symbols
BatchKernelFn pick_batch_x4_kernel(int mode, bool affine);

}  // namespace anyseq
