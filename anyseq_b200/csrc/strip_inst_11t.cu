// strip kernels with LOCAL=1, AFFINE=1 and end-cell tracking (see strip_inst.inl)
#define ANYSEQ_INST_LOCAL true
#define ANYSEQ_INST_AFFINE true
#define ANYSEQ_INST_TRACK 1
#define ANYSEQ_INST_NAME get_strip_kernel_11t
#include "strip_inst.inl"
