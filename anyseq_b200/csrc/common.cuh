// common.cuh -- shared definitions of the B200 alignment engine (sm_100a only).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace anyseq {

constexpr int kWarp = 32;
// "minus infinity" for the affine E/F borders (MIN/2 so that
// adding gap costs cannot wrap).  Scores must stay inside (-2^29, 2^29).
constexpr int kNegInf = -(1 << 30);
// reference SCORE_MIN_VALUE: src/align.impala:16
constexpr int kScoreMin = -2147483647;

enum Mode : int { kGlobal = 0, kSemiglobal = 1, kLocal = 2 };

// status words written by kernels (device -> host)
enum StatusCode : int {
    kStatusOk = 0,
    kStatusTimeout = 1,     // a dependency wait exceeded the watchdog
};

// One DP rectangle ("job"): rows = query symbols q[0..h), columns = subject
// symbols s[0..w).  A single score() call is one job; one Hirschberg level is
// one job per half (forward halves on the forward copies, reversed halves on
// the reversed copies of the sequences).  A job is cut into column strips of
// 32*K columns (one warp each) and row bands of band_h rows; (band, strip)
// pairs are the work items.
//
// Boundary storage mirrors the reference's O(m+n) scheme (column vector, row
// vector, corners: src/scoring.impala:218-259), re-laid for streaming:
//   colH/colE[h]  strip x reads rows from it (written by strip x-1 or the init
//                 kernel) and overwrites them with its own right edge.
//   rowH/rowF[w'] item (b,x) loads its top border from it and stores its
//                 bottom border back (w' = w rounded up to the strip width).
//   corner[x]     H(i0-1, j0-1) for the next band of strip x.
//   progress[x]   number of rows of strip x whose right edge is published
//                 (monotone over bands) -- the only synchronisation.
struct Job {
    const uint8_t* q;
    const uint8_t* s;
    int h, w;
    int band_h;
    int nstrips, nbands;
    long long item_begin;    // prefix sum of nstrips*nbands over jobs
    int* colH;
    int* colE;
    int* rowH;
    int* rowF;
    int* corner;
    int* progress;
    int* best;               // local mode: running maximum (atomicMax)
    // multi-GPU chaining (null on a single GPU): strip 0 reads its left border
    // from inH/inE once *in_progress >= rows; the last strip additionally
    // mirrors its right edge to outH/outE (peer memory) and *out_progress.
    const int* inH;
    const int* inE;
    const int* in_progress;
    int* outH;
    int* outE;
    int* out_progress;
    // initialisation of the borders (init kernel)
    int init_global;         // 1: gap multiples (global), 0: zeros
};

struct ScoreParams {
    int same, diff;
    int gap_extend;          // linear: the gap cost; affine: ge
    int gap_open;            // affine: gi + ge (cost of the first gap symbol); linear: unused
};

#define ANYSEQ_CUDA_CHECK(expr)                                                        \
    do {                                                                               \
        cudaError_t err__ = (expr);                                                    \
        if (err__ != cudaSuccess) {                                                    \
            std::fprintf(stderr, "anyseq_b200: CUDA error %s at %s:%d: %s\n",          \
                         cudaGetErrorName(err__), __FILE__, __LINE__,                  \
                         cudaGetErrorString(err__));                                   \
            return -(int)err__ - 1000;                                                 \
        }                                                                              \
    } while (0)

}  // namespace anyseq
