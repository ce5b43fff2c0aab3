// inbox.cu -- left-border mailboxes for the multi-GPU column-strip wavefront
// (SURVEY.md 8e).  A rank allocates its inbox in its own HBM and exports it as
// a CUDA IPC handle; the producing rank (previous column strip, another
// process on another GPU of the same NVSwitch domain) opens the handle and its
// strip kernel stores its right-edge rows straight into it over NVLink as tagged
// 16-byte records {H, tag, E, tag} (plain st.global on the peer mapping; data and
// tag share an 8-byte word, so no fence and no separate flag are needed).
#include "engine.cuh"

#include <cstring>

namespace anyseq {

int Engine::inbox_create(int rows, Inbox** out, void* handle64)
{
    std::lock_guard<std::recursive_mutex> lock(mu_);
    ANYSEQ_CUDA_CHECK(cudaSetDevice(device));
    Inbox* b = new Inbox();
    b->rows = rows;
    b->owned = true;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, Inbox::bytes_for(rows));
    if (e != cudaSuccess) { delete b; ANYSEQ_CUDA_CHECK(e); }
    b->records = static_cast<int4*>(p);
    e = cudaMemset(p, 0, Inbox::bytes_for(rows));
    if (e != cudaSuccess) { cudaFree(p); delete b; ANYSEQ_CUDA_CHECK(e); }
    if (handle64) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) { cudaFree(p); delete b; ANYSEQ_CUDA_CHECK(e); }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        std::memcpy(handle64, &h, 64);
    }
    *out = b;
    return ANYSEQ_OK;
}

int Engine::inbox_open(const void* handle64, int rows, Inbox** out)
{
    std::lock_guard<std::recursive_mutex> lock(mu_);
    ANYSEQ_CUDA_CHECK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    ANYSEQ_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    Inbox* b = new Inbox();
    b->rows = rows;
    b->owned = false;
    b->records = static_cast<int4*>(p);
    *out = b;
    return ANYSEQ_OK;
}

int Engine::inbox_reset(Inbox* box)
{
    std::lock_guard<std::recursive_mutex> lock(mu_);
    ANYSEQ_CUDA_CHECK(cudaSetDevice(device));
    // records carry a per-run tag, so nothing has to be cleared between runs; a reset
    // only re-synchronises the run counters of the two ends (both must call it)
    box->uses_in = 0;
    box->uses_out = 0;
    if (box->owned) {
        ANYSEQ_CUDA_CHECK(cudaMemset(box->records, 0, Inbox::bytes_for(box->rows)));
        ANYSEQ_CUDA_CHECK(cudaDeviceSynchronize());
    }
    return ANYSEQ_OK;
}

void Engine::inbox_destroy(Inbox* box)
{
    std::lock_guard<std::recursive_mutex> lock(mu_);
    cudaSetDevice(device);
    if (!box->records) return;
    if (box->owned) cudaFree(box->records);
    else cudaIpcCloseMemHandle(box->records);
    box->records = nullptr;
}

}  // namespace anyseq
