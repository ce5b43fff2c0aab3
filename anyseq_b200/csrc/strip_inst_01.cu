// strip kernels with LOCAL=0, AFFINE=1 (see strip_inst.inl)
#define ANYSEQ_INST_LOCAL false
#define ANYSEQ_INST_AFFINE true
#define ANYSEQ_INST_NAME get_strip_kernel_01
#include "strip_inst.inl"
