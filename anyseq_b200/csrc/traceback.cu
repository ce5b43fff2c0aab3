// traceback.cu -- linear-space traceback (placeholder until the GPU path lands).
#include "engine.cuh"
namespace anyseq {
int Engine::align_host(const anyseq_scoring&, const char*, int, const char*, int, char*, char*, anyseq_result*)
{
    set_last_error("anyseq_align: not implemented yet");
    return ANYSEQ_ERR_UNSUPPORTED;
}
}  // namespace anyseq
