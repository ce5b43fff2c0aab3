// batch.cu -- batches of independent short pairs (placeholder).
#include "engine.cuh"
namespace anyseq {
int Engine::score_batch_device(const anyseq_scoring&, const uint8_t*, const int64_t*, const uint8_t*,
                               const int64_t*, int64_t, int32_t*, anyseq_result*)
{
    set_last_error("anyseq_score_batch: not implemented yet");
    return ANYSEQ_ERR_UNSUPPORTED;
}
int Engine::score_batch_host(const anyseq_scoring&, const char*, const int64_t*, const char*, const int64_t*,
                             int64_t, int32_t*, anyseq_result*)
{
    set_last_error("anyseq_score_batch: not implemented yet");
    return ANYSEQ_ERR_UNSUPPORTED;
}
}  // namespace anyseq
