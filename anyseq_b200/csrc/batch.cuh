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
};

using BatchKernelFn = void (*)(const BatchArgs);
// packed 16-bit kernels (batch_x2.cu): two pairs of equal shape per warp; nullptr if K is not instantiated
BatchKernelFn pick_batch_x2_kernel(int mode, bool affine, int K);

}  // namespace anyseq
