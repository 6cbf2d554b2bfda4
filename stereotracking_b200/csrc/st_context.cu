// Per-device library state shared by the entry points (st_common.cuh): SM count, a ring of
// fork / join events and the auxiliary stream paired with each caller stream.  One context per
// CUDA device, created under a mutex at first use on that device, so that two devices in one
// process (or an engine on a non-current device) never share events, streams or opt-ins, and
// concurrent host threads (ctypes releases the GIL) do not race on the ring cursor.
#include <map>
#include <mutex>
#include <new>
#include "st_common.cuh"

struct StDeviceCtx {
    int device = -1;
    int sms = 0;
    std::mutex mu;
    static constexpr int kEvents = 256;
    cudaEvent_t ev[kEvents];
    unsigned ev_next = 0;
    std::map<cudaStream_t, cudaStream_t> aux;
};

namespace {
constexpr int kMaxDevices = 64;
StDeviceCtx *g_ctx[kMaxDevices];
std::mutex g_mu;
}  // namespace

StDeviceCtx *st_ctx() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_ctx[dev]) return g_ctx[dev];
    StDeviceCtx *c = new (std::nothrow) StDeviceCtx();
    if (!c) return nullptr;
    c->device = dev;
    if (cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c->sms <= 0) c->sms = 148;
    for (int i = 0; i < StDeviceCtx::kEvents; ++i)
        if (cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            for (int k = 0; k < i; ++k) cudaEventDestroy(c->ev[k]);
            delete c;
            return nullptr;                       // nothing half-initialised is kept
        }
    g_ctx[dev] = c;
    return c;
}

int st_ctx_sms() {
    StDeviceCtx *c = st_ctx();
    return c ? c->sms : 148;
}

int st_ctx_aux(cudaStream_t caller, cudaStream_t *aux, cudaEvent_t *ev_fork, cudaEvent_t *ev_join) {
    StDeviceCtx *c = st_ctx();
    if (!c) return ST_ERR_CUDA;
    std::lock_guard<std::mutex> lock(c->mu);
    auto it = c->aux.find(caller);
    if (it == c->aux.end()) {
        // one auxiliary stream per caller stream (a shared one would order the forked kernels of
        // independent callers - the two K3 stages of the frame pipeline - behind each other);
        // it inherits the caller's priority.  Streams live as long as the process.
        int prio = 0;
        cudaStream_t s = nullptr;
        if (cudaStreamGetPriority(caller, &prio) != cudaSuccess) { prio = 0; cudaGetLastError(); }
        if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio) != cudaSuccess) return ST_ERR_CUDA;
        it = c->aux.emplace(caller, s).first;
    }
    *aux = it->second;
    *ev_fork = c->ev[c->ev_next++ % StDeviceCtx::kEvents];
    *ev_join = c->ev[c->ev_next++ % StDeviceCtx::kEvents];
    return ST_OK;
}

// Pairs `caller` with an auxiliary stream the caller created itself (the frame pipeline creates its streams
// inside an SM partition and wants K3's forked kernels to stay there).  Must run before the first
// st_ctx_aux(caller, ...); an existing pairing is kept and ST_ERR_INVALID_ARG returned.
int st_ctx_set_aux(cudaStream_t caller, cudaStream_t aux) {
    StDeviceCtx *c = st_ctx();
    if (!c) return ST_ERR_CUDA;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->aux.find(caller) != c->aux.end()) return ST_ERR_INVALID_ARG;
    c->aux.emplace(caller, aux);
    return ST_OK;
}

// Forgets the pairing of `caller` (its streams are about to be destroyed; a later stream may reuse the handle).
void st_ctx_drop_aux(cudaStream_t caller) {
    StDeviceCtx *c = st_ctx();
    if (!c) return;
    std::lock_guard<std::mutex> lock(c->mu);
    c->aux.erase(caller);
}
