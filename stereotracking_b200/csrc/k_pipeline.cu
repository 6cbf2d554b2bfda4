// One lockstep frame of the whole path in a single C call (st_frame_step), and the software
// pipeline over consecutive frames (st_pipeline_*), optionally replayed as CUDA graphs.
// Mirrors OCSORT_Disparity.predict between the detector backbone and the return value
// (mmtrack/models/mot/ocsort_disparity.py:79-109).
#include <cstdlib>
#include <cstring>
#include <new>
#include <cuda.h>
#include "st_common.cuh"

namespace {

// the frame's buffers are consistent with the depth source it asks for
bool io_ok(const st_frame_io *io) {
    if (io->depth_from_raw) return io->raw && io->Hr == io->H && io->Wr == io->W;
    return io->depth_map != nullptr;
}

// K3 on the detections (ocsort_disparity.py:82-83) from the source the frame asks for
int k3_on_dets(st_handle *h, int S, const st_frame_io *io, const StFrameInputs *tab, void *stream) {
    const int cap = h->cfg.max_dets;
    if (io->depth_from_raw)
        return st_box_depth_raw_tab(io->raw, tab, S, io->H, io->W, io->Hp, io->Wp, h->cfg.baseline, h->cfg.focal_length,
                                    io->det_boxes, io->det_counts, cap, io->det_depth, io->det_scales,
                                    io->det_inflated, io->k3_workspace, nullptr, nullptr, nullptr, stream);
    return st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->det_boxes, io->det_counts, cap, io->det_depth,
                        io->det_scales, io->det_inflated, io->k3_workspace, stream);
}

// K3 on the deflated track boxes (ocsort_disparity.py:100-101): rows whose truncated corners
// equal those of their detection take its statistic (st_box_depth_reuse)
int k3_on_tracks(st_handle *h, int S, const st_frame_io *io, const StFrameInputs *tab, void *stream) {
    const int cap = h->cfg.max_dets;
    const float *pb = io->out_src ? io->det_boxes : nullptr, *pd = io->out_src ? io->det_depth : nullptr;
    if (io->depth_from_raw)
        return st_box_depth_raw_tab(io->raw, tab, S, io->H, io->W, io->Hp, io->Wp, h->cfg.baseline, h->cfg.focal_length,
                                    io->out_boxes, io->out_counts, cap, io->out_depth, io->out_scale2, nullptr,
                                    io->k3_workspace, pb, io->out_src, pd, stream);
    return st_box_depth_reuse(io->depth_map, S, io->Hp, io->Wp, io->out_boxes, io->out_counts, cap, io->out_depth,
                              io->out_scale2, nullptr, io->k3_workspace, pb, io->out_src, pd, stream);
}

int track(st_handle *h, int S, const st_frame_io *io, void *stream) {
    return st_track_step(h, S, io->frame_ids, io->det_inflated, io->det_scores, io->det_labels, io->det_scales,
                         io->det_depth, io->det_counts, h->cfg.max_dets, io->trk_boxes, io->out_scores, io->out_labels,
                         io->out_scales, io->trk_depth, io->out_ids, io->out_counts, io->out_src, io->warp, stream);
}

__global__ void k_set_inputs(StFrameInputs *dst, StFrameInputs v) { *dst = v; }

}  // namespace

extern "C" int st_frame_step(st_handle *h, int S, const st_frame_io *io, void *stream) {
    if (!h || !io || S < 0 || S > h->cfg.max_sequences || !io_ok(io)) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    StDeviceGuard guard(h->device);
    const int cap = h->cfg.max_dets;
    int rc;
    if (io->raw && io->depth_map) {
        rc = st_disp2depth(io->raw, S, io->H, io->W, io->Hr, io->Wr, io->Hp, io->Wp, io->depth_map,
                           io->disp3, io->mask, h->cfg.baseline, h->cfg.focal_length, stream);
        if (rc != ST_OK) return rc;
    }
    if (io->cls) {
        rc = st_decode_nms(h, io->cls, io->bbox, io->obj, io->num_levels, io->level_h, io->level_w,
                           io->strides, S, io->num_classes, io->scale_w, io->scale_h, io->ori_h, io->ori_w,
                           cap, io->det_boxes, io->det_scores, io->det_labels, io->det_prior, io->det_keep,
                           io->det_counts, stream);
        if (rc != ST_OK) return rc;
    }
    rc = k3_on_dets(h, S, io, nullptr, stream);
    if (rc != ST_OK) return rc;
    rc = track(h, S, io, stream);
    if (rc != ST_OK) return rc;
    rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, cap, 1, io->out_boxes, stream);
    if (rc != ST_OK) return rc;
    return k3_on_tracks(h, S, io, nullptr, stream);
}


// ---------------------------------------------------------------------------------
// Software pipeline over consecutive frames.  Only the tracker state couples frame
// f+1 to frame f; everything before the tracker (K1, K2, K3 on the detections) and
// after it (deflate, K3 on the tracks) depends on that frame's inputs alone, and K1
// and K2 do not depend on each other.  The latency-bound stages (one CTA per
// sequence or per box, far fewer CTAs than SMs) therefore overlap with each other
// and with the bandwidth-bound K1 of later frames instead of leaving the GPU idle.
// A frame is three segments:
//   pre   on the ring slot's own stream: K2 -> K3 on the detections, with K1 on a side stream
//         beside K2 (joined before K3) when the depth map is written; the pre segments of the
//         frames in flight run concurrently (each slot has its own buffers and K2 scratch)
//   track K4/K5 on the ONE tracker stream, frames strictly in order: the only stage with
//         cross-frame state
//   post  deflate + K3 on the tracks, again on the ring slot's stream (so the post segments of
//         consecutive frames overlap too, and the slot's next frame starts after it by stream order)
// linked by per-slot events.  The tracker stream is the only resource frames queue for.
// Results are identical to st_frame_step: the same kernels run on the same data in the same
// per-frame order.
//
// CUDA graphs (st_pipeline_use_graphs): the pre and the post segment of every ring slot are
// captured once (stream capture of exactly the calls above) and replayed with
// cudaGraphLaunch; the tracker kernel stays a plain launch between them because it carries the
// cross-frame dependency through ordinary stream order.  The frame's INPUT pointers (raw frame,
// head maps) are not baked into the graphs: the kernels read them from a per-slot device table
// (StFrameInputs) that a one-thread kernel updates in stream order when they change, so a
// replay on new inputs costs no graph update.  Per frame the host then issues 2 graph launches,
// 1 kernel launch and 4 event operations instead of ~13 launches and ~25 event operations.
constexpr int kStages = 5;             // trace stages: K1, K2, K3 on dets, tracker, post

struct GraphKey {                      // everything a captured segment bakes in
    st_frame_io io;                    // with the per-frame input pointers zeroed
    int S;
    int lh[ST_MAX_LEVELS], lw[ST_MAX_LEVELS], stride[ST_MAX_LEVELS];
    unsigned align;                    // alignment class of the inputs (selects kernel variants)
};

struct st_pipeline {
    st_handle *h;
    int depth;
    int skip;                          // diagnostics (ST_PIPE_SKIP): bit i = leave stage i's kernels out
    long long submitted;
    cudaStream_t *pre, *side;          // [depth]
    cudaStream_t *aux;                 // [depth] K3's auxiliary stream of every slot when the SMs are partitioned
    cudaStream_t trk;
    // SM partition (green contexts, ST_PIPE_PARTITION): the tracker stream owns trk_sms SMs, every other
    // stream of the pipeline the rest, so that the latency-bound tracker CTAs never share an SM with the
    // streaming stages of the neighbouring frames
    void *green_trk, *green_rest;
    int trk_sms, rest_sms;
    cudaEvent_t *ev;                   // [4][depth]: pre done, tracker done, K1 fork, K1 join
    cudaEvent_t fork_ev;
    void *scratch_mem;                 // K2 candidate scratch of every slot + the input tables
    StK2Scratch *scratch;              // [depth]
    StFrameInputs *tab_dev;            // [depth] device
    StFrameInputs *tab_host;           // [depth] what each slot's table holds
    bool *tab_valid;
    int use_graphs;
    cudaGraphExec_t *g_pre, *g_post;   // [depth]
    bool *have_graph;
    GraphKey *key;                     // [depth]
    long long n_graph_frames, n_captures;
    // optional timeline (st_pipeline_trace): begin / end timing events of every stage of the
    // first trace_cap frames submitted after tracing was switched on (a frame replayed as graphs
    // is stamped per segment: stages 0-2 all show the whole pre segment)
    cudaEvent_t *trace_ev;             // [trace_cap][kStages][2]
    cudaEvent_t trace_origin;
    int trace_cap, trace_n;
};

namespace {

enum { EV_PRE = 0, EV_TRK = 1, EV_K1F = 2, EV_K1J = 3, EV_KINDS = 4 };

void fill_inputs(const st_frame_io *io, StFrameInputs *t) {
    memset(t, 0, sizeof(*t));
    t->raw = io->raw;
    if (io->cls)
        for (int l = 0; l < io->num_levels && l < ST_MAX_LEVELS; ++l) {
            t->cls[l] = io->cls[l]; t->bbox[l] = io->bbox[l]; t->obj[l] = io->obj[l];
        }
}

void fill_key(const st_frame_io *io, int S, GraphKey *k) {
    memset(k, 0, sizeof(*k));
    k->io = *io;
    k->io.frame_ids = nullptr; k->io.raw = nullptr; k->io.warp = nullptr;
    k->io.cls = nullptr; k->io.bbox = nullptr; k->io.obj = nullptr;
    k->io.level_h = nullptr; k->io.level_w = nullptr; k->io.strides = nullptr;
    k->S = S;
    unsigned a = (unsigned)((uintptr_t)io->raw & 15u) | (io->raw ? 16u : 0u) | (io->cls ? 32u : 0u);
    if (io->cls)
        for (int l = 0; l < io->num_levels && l < ST_MAX_LEVELS; ++l) {
            k->lh[l] = io->level_h[l]; k->lw[l] = io->level_w[l]; k->stride[l] = io->strides[l];
            a = a * 31u + (unsigned)(((uintptr_t)io->cls[l] & 15u) | (((uintptr_t)io->obj[l] & 15u) << 4) |
                                     (((uintptr_t)io->bbox[l] & 15u) << 8));
        }
    k->align = a;
}

// pre segment on the slot's stream (and its side stream): everything of the frame that does not
// depend on the tracker.  `traced`: record the timeline events of stages 0-2.
int seg_pre(st_pipeline *p, int slot, int S, const st_frame_io *io, bool traced) {
    st_handle *h = p->h;
    const int cap = h->cfg.max_dets;
    cudaStream_t pre = p->pre[slot], side = p->side[slot];
    const StFrameInputs *tab = p->tab_dev + slot;
    auto ev = [&](int kind) { return p->ev[kind * p->depth + slot]; };
    auto mark = [&](int stage, int end, cudaStream_t st) {
        if (traced) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + stage) * 2 + end], st);
    };
    const int skip = p->skip;
    int rc;
    const bool k1 = io->raw && io->depth_map;
    if (k1) {
        ST_CUDA_OK(cudaEventRecord(ev(EV_K1F), pre));
        ST_CUDA_OK(cudaStreamWaitEvent(side, ev(EV_K1F), 0));
        mark(0, 0, side);
        if (!(skip & 1)) {
            rc = st_disp2depth_tab(io->raw, tab, S, io->H, io->W, io->Hr, io->Wr, io->Hp, io->Wp, io->depth_map,
                                   io->disp3, io->mask, h->cfg.baseline, h->cfg.focal_length, side);
            if (rc != ST_OK) return rc;
        }
        mark(0, 1, side);
        ST_CUDA_OK(cudaEventRecord(ev(EV_K1J), side));
    } else {
        mark(0, 0, pre); mark(0, 1, pre);
    }
    mark(1, 0, pre);
    if (io->cls && !(skip & 2)) {
        rc = st_decode_nms_ws(h, io->cls, io->bbox, io->obj, tab, p->scratch + slot, io->num_levels, io->level_h,
                              io->level_w, io->strides, S, io->num_classes, io->scale_w, io->scale_h, io->ori_h,
                              io->ori_w, cap, io->det_boxes, io->det_scores, io->det_labels, io->det_prior,
                              io->det_keep, io->det_counts, pre);
        if (rc != ST_OK) return rc;
    }
    mark(1, 1, pre);
    if (k1) ST_CUDA_OK(cudaStreamWaitEvent(pre, ev(EV_K1J), 0));
    mark(2, 0, pre);
    rc = (skip & 4) ? ST_OK : k3_on_dets(h, S, io, tab, pre);
    if (rc != ST_OK) return rc;
    mark(2, 1, pre);
    return ST_OK;
}

int seg_post(st_pipeline *p, int slot, int S, const st_frame_io *io) {
    st_handle *h = p->h;
    cudaStream_t st = p->pre[slot];
    int rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, h->cfg.max_dets, 1, io->out_boxes, st);
    if (rc != ST_OK) return rc;
    return k3_on_tracks(h, S, io, p->tab_dev + slot, st);
}

void drop_graphs(st_pipeline *p, int slot) {
    if (p->have_graph[slot]) {
        cudaGraphExecDestroy(p->g_pre[slot]);
        cudaGraphExecDestroy(p->g_post[slot]);
        p->have_graph[slot] = false;
    }
}

// capture the two segments of `slot` for this frame shape (stream capture of the very calls the
// uncaptured path makes)
int capture(st_pipeline *p, int slot, int S, const st_frame_io *io) {
    drop_graphs(p, slot);
    cudaGraph_t g = nullptr;
    ST_CUDA_OK(cudaStreamBeginCapture(p->pre[slot], cudaStreamCaptureModeRelaxed));
    int rc = seg_pre(p, slot, S, io, false);
    cudaError_t e = cudaStreamEndCapture(p->pre[slot], &g);
    if (rc != ST_OK || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc != ST_OK ? rc : ST_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&p->g_pre[slot], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return ST_ERR_CUDA;
    g = nullptr;
    if (cudaStreamBeginCapture(p->pre[slot], cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGraphExecDestroy(p->g_pre[slot]);
        return ST_ERR_CUDA;
    }
    rc = seg_post(p, slot, S, io);
    e = cudaStreamEndCapture(p->pre[slot], &g);
    if (rc == ST_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&p->g_post[slot], g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc != ST_OK || e != cudaSuccess) {
        cudaGraphExecDestroy(p->g_pre[slot]);
        cudaGetLastError();
        return rc != ST_OK ? rc : ST_ERR_CUDA;
    }
    p->have_graph[slot] = true;
    p->n_captures += 1;
    return ST_OK;
}

void free_trace(st_pipeline *p) {
    if (!p->trace_ev) return;
    for (int i = 0; i < p->trace_cap * kStages * 2; ++i) cudaEventDestroy(p->trace_ev[i]);
    cudaEventDestroy(p->trace_origin);
    delete[] p->trace_ev;
    p->trace_ev = nullptr;
}

// ---- SM partition through green contexts (driver entry points fetched at run time: no libcuda link) ----
struct GreenApi {
    decltype(&cuDeviceGet) deviceGet;
    decltype(&cuDeviceGetDevResource) getDevResource;
    decltype(&cuDevSmResourceSplitByCount) splitByCount;
    decltype(&cuDevResourceGenerateDesc) generateDesc;
    decltype(&cuGreenCtxCreate) ctxCreate;
    decltype(&cuGreenCtxDestroy) ctxDestroy;
    decltype(&cuGreenCtxStreamCreate) streamCreate;
    bool ok;
};

const GreenApi &green_api() {
    static const GreenApi api = [] {
        GreenApi a;
        memset(&a, 0, sizeof(a));
        auto get = [](const char *name, void **f) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *f;
        };
        a.ok = get("cuDeviceGet", (void **)&a.deviceGet) && get("cuDeviceGetDevResource", (void **)&a.getDevResource) &&
               get("cuDevSmResourceSplitByCount", (void **)&a.splitByCount) &&
               get("cuDevResourceGenerateDesc", (void **)&a.generateDesc) && get("cuGreenCtxCreate", (void **)&a.ctxCreate) &&
               get("cuGreenCtxDestroy", (void **)&a.ctxDestroy) && get("cuGreenCtxStreamCreate", (void **)&a.streamCreate);
        cudaGetLastError();
        return a;
    }();
    return api;
}

// two green contexts on `device`: at least `want` SMs for the tracker, the remainder for everything else
bool partition_create(int device, int want, st_pipeline *p) {
    const GreenApi &a = green_api();
    if (!a.ok) return false;
    CUdevice dev;
    CUdevResource all, part, rest;
    unsigned groups = 1;
    if (a.deviceGet(&dev, device) != CUDA_SUCCESS || a.getDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS ||
        a.splitByCount(&part, &groups, &all, &rest, 0, (unsigned)want) != CUDA_SUCCESS || groups != 1 || rest.sm.smCount < 16)
        return false;
    CUdevResourceDesc d_trk, d_rest;
    CUgreenCtx g_trk = nullptr, g_rest = nullptr;
    if (a.generateDesc(&d_trk, &part, 1) != CUDA_SUCCESS || a.generateDesc(&d_rest, &rest, 1) != CUDA_SUCCESS ||
        a.ctxCreate(&g_trk, d_trk, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS)
        return false;
    if (a.ctxCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { a.ctxDestroy(g_trk); return false; }
    p->green_trk = g_trk; p->green_rest = g_rest;
    p->trk_sms = (int)part.sm.smCount; p->rest_sms = (int)rest.sm.smCount;
    return true;
}

bool partition_stream(void *green, int priority, cudaStream_t *s) {
    CUstream cs = nullptr;
    if (green_api().streamCreate(&cs, (CUgreenCtx)green, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) return false;
    *s = (cudaStream_t)cs;
    return true;
}

}  // namespace

extern "C" int st_pipeline_create(st_handle *h, int depth, st_pipeline **out) {
    if (!h || !out || depth < 1 || depth > 64) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    st_pipeline *p = new (std::nothrow) st_pipeline();
    if (!p) return ST_ERR_ALLOC;
    memset(p, 0, sizeof(*p));
    p->h = h; p->depth = depth;
    { const char *e = getenv("ST_PIPE_SKIP"); p->skip = e ? atoi(e) : 0; }
    p->pre = new (std::nothrow) cudaStream_t[depth]();
    p->side = new (std::nothrow) cudaStream_t[depth]();
    p->aux = new (std::nothrow) cudaStream_t[depth]();
    p->ev = new (std::nothrow) cudaEvent_t[EV_KINDS * depth]();
    p->scratch = new (std::nothrow) StK2Scratch[depth]();
    p->tab_host = new (std::nothrow) StFrameInputs[depth]();
    p->tab_valid = new (std::nothrow) bool[depth]();
    p->g_pre = new (std::nothrow) cudaGraphExec_t[depth]();
    p->g_post = new (std::nothrow) cudaGraphExec_t[depth]();
    p->have_graph = new (std::nothrow) bool[depth]();
    p->key = new (std::nothrow) GraphKey[depth]();
    bool ok = p->pre && p->side && p->aux && p->ev && p->scratch && p->tab_host && p->tab_valid && p->g_pre && p->g_post &&
              p->have_graph && p->key;
    ok = ok && cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    // Stream priorities: K1 (bandwidth-bound, fills every SM for its whole duration) on the lowest
    // priority, the latency-bound stages on the highest, so that their CTAs are placed first
    // whenever slots free up (+5 % frames/s at the bench workload; ST_PIPE_PRIO=0 switches it off).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char *pe = getenv("ST_PIPE_PRIO");
    const int prio_mode = pe ? atoi(pe) : 2;
    const int hi = prio_mode ? prio_hi : 0, lo = prio_mode ? prio_lo : 0;
    // SM partition: one tracker CTA per sequence, one SM each; worth it while that leaves most of the device to
    // the other stages (ST_PIPE_PARTITION=0 / 1 forces it off / on)
    {
        const char *ge = getenv("ST_PIPE_PARTITION");
        const int want = (h->cfg.max_sequences + 7) & ~7;
        const bool wanted = ge ? atoi(ge) != 0 : false;
        if (ok && wanted && want * 20 <= st_ctx_sms() * 11) partition_create(h->device, want, p);
    }
    for (int i = 0; ok && i < depth; ++i) {
        if (p->green_rest)
            ok = partition_stream(p->green_rest, hi, &p->pre[i]) && partition_stream(p->green_rest, lo, &p->side[i]) &&
                 partition_stream(p->green_rest, hi, &p->aux[i]) && st_ctx_set_aux(p->pre[i], p->aux[i]) == ST_OK;
        else
            ok = cudaStreamCreateWithPriority(&p->pre[i], cudaStreamNonBlocking, hi) == cudaSuccess &&
                 cudaStreamCreateWithPriority(&p->side[i], cudaStreamNonBlocking, lo) == cudaSuccess;
    }
    if (p->green_trk) ok = ok && partition_stream(p->green_trk, hi, &p->trk);
    else ok = ok && cudaStreamCreateWithPriority(&p->trk, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int i = 0; ok && i < EV_KINDS * depth; ++i)
        ok = cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming) == cudaSuccess;
    // K2 candidate scratch per slot (the pre segments of frames in flight run concurrently) and
    // the input tables, one allocation
    const size_t S = (size_t)h->cfg.max_sequences, N = (size_t)h->cfg.max_candidates;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t per_slot = up(sizeof(int32_t) * S) + 3 * up(sizeof(int32_t) * S * N);
    const size_t tab_bytes = up(sizeof(StFrameInputs) * (size_t)depth);
    if (ok) ok = cudaMalloc(&p->scratch_mem, tab_bytes + per_slot * depth) == cudaSuccess;
    if (ok) {
        unsigned char *c = static_cast<unsigned char *>(p->scratch_mem);
        p->tab_dev = reinterpret_cast<StFrameInputs *>(c);
        c += tab_bytes;
        for (int i = 0; i < depth; ++i) {
            p->scratch[i].count = reinterpret_cast<int32_t *>(c); c += up(sizeof(int32_t) * S);
            p->scratch[i].score = reinterpret_cast<float *>(c); c += up(sizeof(float) * S * N);
            p->scratch[i].prior = reinterpret_cast<int32_t *>(c); c += up(sizeof(int32_t) * S * N);
            p->scratch[i].label = reinterpret_cast<int32_t *>(c); c += up(sizeof(int32_t) * S * N);
        }
        // the auxiliary streams K3 forks onto exist before anything is captured
        cudaStream_t aux; cudaEvent_t e1, e2;
        for (int i = 0; ok && i < depth; ++i) ok = st_ctx_aux(p->pre[i], &aux, &e1, &e2) == ST_OK;
    }
    if (!ok) { st_pipeline_destroy(p); return ST_ERR_CUDA; }
    *out = p;
    return ST_OK;
}

extern "C" int st_pipeline_destroy(st_pipeline *p) {
    if (!p) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(p->h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < p->depth; ++i) {
        if (p->have_graph) drop_graphs(p, i);
        if (p->aux && p->aux[i]) {                       // partitioned: the pairing dies with the streams
            if (p->pre && p->pre[i]) st_ctx_drop_aux(p->pre[i]);
            cudaStreamDestroy(p->aux[i]);
        }
        if (p->pre && p->pre[i]) cudaStreamDestroy(p->pre[i]);
        if (p->side && p->side[i]) cudaStreamDestroy(p->side[i]);
    }
    if (p->ev)
        for (int i = 0; i < EV_KINDS * p->depth; ++i)
            if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    if (p->fork_ev) cudaEventDestroy(p->fork_ev);
    if (p->trk) cudaStreamDestroy(p->trk);
    if (p->green_trk) green_api().ctxDestroy((CUgreenCtx)p->green_trk);
    if (p->green_rest) green_api().ctxDestroy((CUgreenCtx)p->green_rest);
    free_trace(p);
    if (p->scratch_mem) cudaFree(p->scratch_mem);
    delete[] p->pre; delete[] p->side; delete[] p->aux; delete[] p->ev; delete[] p->scratch; delete[] p->tab_host;
    delete[] p->tab_valid; delete[] p->g_pre; delete[] p->g_post; delete[] p->have_graph; delete[] p->key;
    delete p;
    return ST_OK;
}

extern "C" int st_pipeline_use_graphs(st_pipeline *p, int enable) {
    if (!p) return ST_ERR_INVALID_ARG;
    p->use_graphs = enable != 0;
    return ST_OK;
}

extern "C" int st_pipeline_partition(st_pipeline *p, int *tracker_sms, int *other_sms) {
    if (!p || !tracker_sms || !other_sms) return ST_ERR_INVALID_ARG;
    *tracker_sms = p->green_trk ? p->trk_sms : 0;
    *other_sms = p->green_trk ? p->rest_sms : 0;
    return ST_OK;
}

extern "C" int st_pipeline_stats(st_pipeline *p, long long *graph_frames, long long *captures) {
    if (!p) return ST_ERR_INVALID_ARG;
    if (graph_frames) *graph_frames = p->n_graph_frames;
    if (captures) *captures = p->n_captures;
    return ST_OK;
}

extern "C" int st_pipeline_fork(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(p->h->device);
    ST_CUDA_OK(cudaEventRecord(p->fork_ev, (cudaStream_t)stream));
    for (int i = 0; i < p->depth; ++i) ST_CUDA_OK(cudaStreamWaitEvent(p->pre[i], p->fork_ev, 0));
    ST_CUDA_OK(cudaStreamWaitEvent(p->trk, p->fork_ev, 0));
    return ST_OK;
}

extern "C" int st_pipeline_join(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(p->h->device);
    auto join = [&](cudaStream_t s) -> int {
        ST_CUDA_OK(cudaEventRecord(p->fork_ev, s));
        ST_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, p->fork_ev, 0));
        return ST_OK;
    };
    for (int i = 0; i < p->depth; ++i) {
        if (join(p->pre[i]) != ST_OK) return ST_ERR_CUDA;
        if (join(p->side[i]) != ST_OK) return ST_ERR_CUDA;
    }
    if (join(p->trk) != ST_OK) return ST_ERR_CUDA;
    return ST_OK;
}

extern "C" int st_pipeline_submit(st_pipeline *p, int S, const st_frame_io *io, void *inputs_ready,
                                  void *frame_done) {
    if (!p || !io || S < 0 || S > p->h->cfg.max_sequences || !io_ok(io)) return ST_ERR_INVALID_ARG;
    if (io->cls && (io->num_levels < 1 || io->num_levels > ST_MAX_LEVELS)) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    st_handle *h = p->h;
    StDeviceGuard guard(h->device);
    const int slot = (int)(p->submitted % p->depth);
    auto ev = [&](int kind) { return p->ev[kind * p->depth + slot]; };
    const bool reused = p->submitted >= p->depth;
    const bool traced = p->trace_ev && p->trace_n < p->trace_cap;
    cudaStream_t pre = p->pre[slot];
    int rc;
    // (the buffer set of this slot is free: its previous frame's post segment ran on this stream)
    if (inputs_ready) ST_CUDA_OK(cudaStreamWaitEvent(pre, (cudaEvent_t)inputs_ready, 0));
    // the slot's input table follows the frame's input pointers (stream-ordered, parameters by value)
    StFrameInputs tab;
    fill_inputs(io, &tab);
    if (!p->tab_valid[slot] || memcmp(&tab, &p->tab_host[slot], sizeof(tab)) != 0) {
        k_set_inputs<<<1, 1, 0, pre>>>(p->tab_dev + slot, tab);
        p->tab_host[slot] = tab;
        p->tab_valid[slot] = true;
    }
    // graphs from the slot's second use on: the first pass warms up every lazy initialisation
    // (opt-ins, occupancy queries, auxiliary streams) outside of a capture
    bool graph = p->use_graphs && reused && p->skip == 0;
    if (graph) {
        GraphKey k;
        fill_key(io, S, &k);
        if (!p->have_graph[slot] || memcmp(&k, &p->key[slot], sizeof(k)) != 0) {
            rc = capture(p, slot, S, io);
            if (rc != ST_OK) return rc;
            p->key[slot] = k;
        }
    }
    // ---- pre: K1 | K2 -> K3 on the detections
    if (graph) {
        // a replayed segment is time-stamped as a whole: stages 0 and 1 carry its begin, stage 2 its end
        if (traced)
            for (int st_ = 0; st_ < 3; ++st_) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + st_) * 2 + 0], pre);
        ST_CUDA_OK(cudaGraphLaunch(p->g_pre[slot], pre));
        if (traced)
            for (int st_ = 0; st_ < 3; ++st_) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + st_) * 2 + 1], pre);
    } else if ((rc = seg_pre(p, slot, S, io, traced)) != ST_OK) return rc;
    ST_CUDA_OK(cudaEventRecord(ev(EV_PRE), pre));
    // ---- the tracker, frames strictly in order on one stream
    ST_CUDA_OK(cudaStreamWaitEvent(p->trk, ev(EV_PRE), 0));
    if (traced) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + 3) * 2 + 0], p->trk);
    rc = track(h, S, io, p->trk);
    if (rc != ST_OK) return rc;
    if (traced) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + 3) * 2 + 1], p->trk);
    ST_CUDA_OK(cudaEventRecord(ev(EV_TRK), p->trk));
    // ---- post: deflate and the output depth, back on the slot's stream
    ST_CUDA_OK(cudaStreamWaitEvent(pre, ev(EV_TRK), 0));
    if (traced) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + 4) * 2 + 0], pre);
    if (graph) ST_CUDA_OK(cudaGraphLaunch(p->g_post[slot], pre));
    else if ((rc = seg_post(p, slot, S, io)) != ST_OK) return rc;
    if (traced) { cudaEventRecord(p->trace_ev[(p->trace_n * kStages + 4) * 2 + 1], pre); p->trace_n += 1; }
    if (frame_done) ST_CUDA_OK(cudaEventRecord((cudaEvent_t)frame_done, pre));
    if (graph) p->n_graph_frames += 1;
    p->submitted += 1;
    return ST_OK;
}

// Timeline of the pipeline (tracing aid): after st_pipeline_trace(p, n) the begin and end
// of every stage of the next n submitted frames are time-stamped with CUDA events on the
// stage's own stream; st_pipeline_read_trace synchronises and returns, per traced frame
// and stage, {begin, end} in milliseconds since tracing was switched on.
extern "C" int st_pipeline_trace(st_pipeline *p, int n_frames) {
    if (!p || n_frames < 0 || n_frames > 4096) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(p->h->device);
    free_trace(p);
    p->trace_cap = n_frames; p->trace_n = 0;
    if (n_frames == 0) return ST_OK;
    p->trace_ev = new (std::nothrow) cudaEvent_t[(size_t)n_frames * kStages * 2];
    if (!p->trace_ev) { p->trace_cap = 0; return ST_ERR_ALLOC; }
    for (int i = 0; i < n_frames * kStages * 2; ++i) ST_CUDA_OK(cudaEventCreate(&p->trace_ev[i]));
    ST_CUDA_OK(cudaEventCreate(&p->trace_origin));
    ST_CUDA_OK(cudaEventRecord(p->trace_origin, p->trk));
    return ST_OK;
}

extern "C" int st_pipeline_read_trace(st_pipeline *p, float *ms_host, int max_frames) {
    if (!p || !ms_host || max_frames < 0) return ST_ERR_INVALID_ARG;
    if (!p->trace_ev) return 0;
    StDeviceGuard guard(p->h->device);
    ST_CUDA_OK(cudaDeviceSynchronize());
    const int n = p->trace_n < max_frames ? p->trace_n : max_frames;
    for (int i = 0; i < n * kStages * 2; ++i)
        ST_CUDA_OK(cudaEventElapsedTime(&ms_host[i], p->trace_origin, p->trace_ev[i]));
    return n;
}
