// strip_inst.inl -- explicit instantiation of the strip kernels of one
// (LOCAL, AFFINE) pair; included by strip_inst_XY.cu with ANYSEQ_INST_LOCAL,
// ANYSEQ_INST_AFFINE and ANYSEQ_INST_NAME defined (4 translation units so the
// 32 kernels compile in parallel).
#include "strip_kernel.cuh"

namespace anyseq {

#ifdef ANYSEQ_INST_TRACK
StripKernelFn ANYSEQ_INST_NAME(int K, bool mask)
#else
StripKernelFn ANYSEQ_INST_NAME(int K, bool mask, int form)
#endif
{
    constexpr bool L = ANYSEQ_INST_LOCAL;
    constexpr bool A = ANYSEQ_INST_AFFINE;
#ifdef ANYSEQ_INST_TRACK
    if (mask) {
        switch (K) {
            case 4: return strip_kernel<L, A, 4, true, true>;
            case 8: return strip_kernel<L, A, 8, true, true>;
            case 16: return strip_kernel<L, A, 16, true, true>;
            case 32: return strip_kernel<L, A, 32, true, true>;
            default: return nullptr;
        }
    }
    switch (K) {
        case 4: return strip_kernel<L, A, 4, false, true>;
        case 8: return strip_kernel<L, A, 8, false, true>;
        case 16: return strip_kernel<L, A, 16, false, true>;
        default: return nullptr;
    }
#else
#if ANYSEQ_INST_AFFINE
    if (mask && form == 2) {          // mixed cells (even rows coupled, odd rows decoupled): two-row tiles only
        switch (K) {
            case 16: return strip_kernel<L, A, 16, true, false, 2>;
            case 32: return strip_kernel<L, A, 32, true, false, 2>;
            default: break;
        }
    }
    if (mask && form == 0) {          // coupled cells: Gotoh, small alphabets (the full-width workhorse)
        switch (K) {
            case 8: return strip_kernel<L, A, 8, true, false, 0>;
            case 16: return strip_kernel<L, A, 16, true, false, 0>;
            case 32: return strip_kernel<L, A, 32, true, false, 0>;
            default: break;
        }
    }
#else
    (void)form;
#endif
    if (mask) {
        switch (K) {
            case 4: return strip_kernel<L, A, 4, true>;
            case 8: return strip_kernel<L, A, 8, true>;
            case 16: return strip_kernel<L, A, 16, true>;
            case 32: return strip_kernel<L, A, 32, true>;
            default: return nullptr;
        }
    }
    switch (K) {
        case 4: return strip_kernel<L, A, 4, false>;
        case 8: return strip_kernel<L, A, 8, false>;
        case 16: return strip_kernel<L, A, 16, false>;
        case 32: return strip_kernel<L, A, 32, false>;
        default: return nullptr;
    }
#endif
}

}  // namespace anyseq
