// strip kernels with LOCAL=1, AFFINE=0 (see strip_inst.inl)
#define ANYSEQ_INST_LOCAL true
#define ANYSEQ_INST_AFFINE false
#define ANYSEQ_INST_NAME get_strip_kernel_10
#include "strip_inst.inl"
