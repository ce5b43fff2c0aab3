// strip kernels with LOCAL=0, AFFINE=0 (see strip_inst.inl)
#define ANYSEQ_INST_LOCAL false
#define ANYSEQ_INST_AFFINE false
#define ANYSEQ_INST_NAME get_strip_kernel_00
#include "strip_inst.inl"
