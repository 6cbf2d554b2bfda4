// K2: YOLOX head post-process for S frames in lockstep, two kernels:
//   k2_score_compact  prior grid, sigmoid(cls)*sigmoid(obj), max over classes,
//                     score > thr, warp-aggregated compaction          (grid over priors)
//   k2_sort_nms       one CTA per sequence, everything in shared memory: stable rank
//                     by (score desc, prior index asc), box decode of the candidates
//                     only, rescale, greedy NMS in 64-box chunks (warp-ballot bitmask
//                     for the diagonal block, kept boxes of the chunk against all later
//                     boxes), clamp, emit in descending score order
// Replaces the third-party post-process the reference calls at
// mmtrack/models/detectors/yolo_detector_disparity_v1.py:121-122 (mmyolo 0.2.0
// YOLOv5Head.predict_by_feat / YOLOXBBoxCoder.decode, mmdet 3.0.0rc4
// filter_scores_and_topk / _bbox_post_process, mmcv 2.0.0rc3 batched_nms).
// (Round-1 history: rank/decode, a global N x N/64 suppression mask and the keep scan
// were three more launches, 95 us of mostly launch and dependent-load latency for
// ~200 candidates per frame.)
#include <cmath>
#include <cstdlib>
#include "st_common.cuh"

namespace {

struct HeadLevels {
    const float *cls[ST_MAX_LEVELS];
    const float *bbox[ST_MAX_LEVELS];
    const float *obj[ST_MAX_LEVELS];
    int h[ST_MAX_LEVELS], w[ST_MAX_LEVELS], stride[ST_MAX_LEVELS];
    int prior_off[ST_MAX_LEVELS + 1];
    int n;
};

// head-map pointers of level l: from the frame's input table when the launch carries one (a
// replayed CUDA graph), else the ones passed with the launch
#define K2_CLS(l) (tab ? tab->cls[l] : L.cls[l])
#define K2_BBOX(l) (tab ? tab->bbox[l] : L.bbox[l])
#define K2_OBJ(l) (tab ? tab->obj[l] : L.obj[l])

__device__ __forceinline__ int level_of(const HeadLevels &L, int p) {
    int l = 0;
    while (l + 1 < L.n && p >= L.prior_off[l + 1]) ++l;
    return l;
}

// float -> uint with the same ordering (handles negatives)
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(256)
k2_score_compact(HeadLevels L, const StFrameInputs *__restrict__ tab, int P, int num_classes, float score_thr, float logit_cut, int cap,
                 int32_t *__restrict__ cand_count, float *__restrict__ cand_score,
                 int32_t *__restrict__ cand_prior, int32_t *__restrict__ cand_label) {
    const int s = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    float score = 0.f;
    int label = 0;
    if (p < P) {
        const int l = level_of(L, p);
        const int q = p - L.prior_off[l];
        const long long hw = (long long)L.h[l] * L.w[l];
        // sigmoid(x) <= score_thr for x <= logit_cut and the other factor is <= 1, so a
        // logit below the cut cannot pass `score > score_thr`: the float64 exp behind the
        // correctly rounded sigmoid is only evaluated for the ~1 % of priors that may
        const float lo = K2_OBJ(l)[(long long)s * hw + q];
        if (lo > logit_cut) {
            const float so = st_sigmoid_f32(lo);
            // scores = sigmoid(cls) * sigmoid(obj); max over classes, first max wins
            float best = -1.f;
            for (int c = 0; c < num_classes; ++c) {
                const float lc = K2_CLS(l)[((long long)s * num_classes + c) * hw + q];
                if (!(lc > logit_cut)) continue;
                float sc = __fmul_rn(st_sigmoid_f32(lc), so);
                if (sc > best) { best = sc; label = c; }
            }
            score = best;
            pass = score > score_thr;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(ballot) - 1)) base = atomicAdd(&cand_count[s], __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    if (pass) {
        int pos = base + __popc(ballot & ((1u << lane) - 1u));
        if (pos < cap) {
            cand_score[(long long)s * cap + pos] = score;
            cand_prior[(long long)s * cap + pos] = p;
            cand_label[(long long)s * cap + pos] = label;
        }
    }
}

// Vectorised form (every level's h*w a multiple of 4, 16-byte aligned maps): the objectness logits
// are streamed with 128-bit loads, and the class logits of a word are only fetched when one of its
// objectness logits is above the cut (~1 % of the words), so the pass reads 4 bytes per prior
// instead of 8.
// The exact score needs float64 exp (correctly rounded sigmoid).  Evaluated where the logits are
// found it ran at 1/32 lane efficiency in two of three warps (14 us for 5 MB); so the streaming
// phase only screens - logit cut, then the product of fast-math sigmoids against 0.9 * thr (their
// error is < 1e-5 relative) - and queues the survivors in shared memory, and the warp then scores
// the queue one candidate per lane.
// (History: a tiled form - a work item = 4 priors of one level of one sequence, per-item level and
// sequence arithmetic, 2 loads per thread - reached 0.41-0.47 of the copy peak at 512-2048 x 1080p,
// its persistent double-buffered variant less; profiles/r02_k2_score_variants.txt.)
constexpr int kScoreWarps = 8;

// The objectness map of a level is ONE contiguous array [S, h*w] over
// all sequences, so the streaming part is a flat walk over S*h*w/4 128-bit words per level with no
// per-item level / sequence arithmetic: a tile is 256 * kItems consecutive words of one level
// (block-uniform pointer and level), a thread issues its kItems loads, takes the maximum of each
// word and leaves when no lane of its warp saw a logit above the cut - the case for nearly every
// warp.  Only warps with a hit compute (sequence, prior) = divmod(element, h*w), fetch the class
// logits, screen and score as above.  A warp's words may straddle two sequences: candidates are
// appended with one atomic per (warp, sequence) via match_any.
struct FlatTiles { int first[ST_MAX_LEVELS + 1]; };

template <int kItems>
__global__ void __launch_bounds__(kScoreWarps * 32)
k2_score_flat(HeadLevels L, const StFrameInputs *__restrict__ tab, FlatTiles ft, int S, int num_classes, float score_thr, float logit_cut,
              int cap, int32_t *__restrict__ cand_count, float *__restrict__ cand_score, int32_t *__restrict__ cand_prior,
              int32_t *__restrict__ cand_label) {
    __shared__ int32_t queue_all[kScoreWarps][128 * kItems];
    const int lane = threadIdx.x & 31;
    int l = 0;
    while (l + 1 < L.n && (int)blockIdx.x >= ft.first[l + 1]) ++l;          // block-uniform
    const int hw = L.h[l] * L.w[l];
    const int n4 = (int)(((long long)S * hw) >> 2);
    const float *__restrict__ obj = K2_OBJ(l);
    const int i0 = ((int)blockIdx.x - ft.first[l]) * (kScoreWarps * 32 * kItems) + threadIdx.x;
    float4 o4s[kItems];
    bool hit = false;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
        const int i = i0 + u * (kScoreWarps * 32);
        o4s[u] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
        if (i < n4)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(o4s[u].x), "=f"(o4s[u].y), "=f"(o4s[u].z), "=f"(o4s[u].w)
                         : "l"(obj + (long long)i * 4));
    }
#pragma unroll
    for (int u = 0; u < kItems; ++u)       // fmaxf drops NaN: a NaN logit is never above the cut, as in the exact test
        hit = hit || fmaxf(fmaxf(o4s[u].x, o4s[u].y), fmaxf(o4s[u].z, o4s[u].w)) > logit_cut;
    if (!__any_sync(0xffffffffu, hit)) return;
    // ---- a lane of this warp saw a logit above the cut (everything below is fully unrolled: the loaded
    // words stay in registers) ----
    int32_t *queue = queue_all[threadIdx.x >> 5];
    const float screen_thr = __fmul_rn(score_thr, 0.9f);
    const float *__restrict__ cls = K2_CLS(l);
    const int p_off = L.prior_off[l];
    // screen: bit 4u+k = prior k of word u survives the logit cut and the fast-math product test
    unsigned keep = 0u;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
        const float4 o4 = o4s[u];
        const float lo[4] = {o4.x, o4.y, o4.z, o4.w};
        if (lo[0] > logit_cut || lo[1] > logit_cut || lo[2] > logit_cut || lo[3] > logit_cut) {
            const int e0 = (i0 + u * (kScoreWarps * 32)) * 4;               // element of the level's [S, hw] array
            const int s = e0 / hw, q = e0 - s * hw;                         // hw % 4 == 0: the word lies in one sequence
            float mc[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};         // max class logit
            for (int c = 0; c < num_classes; ++c) {
                const float4 c4 = *reinterpret_cast<const float4 *>(cls + ((long long)s * num_classes + c) * hw + q);
                mc[0] = fmaxf(mc[0], c4.x); mc[1] = fmaxf(mc[1], c4.y);
                mc[2] = fmaxf(mc[2], c4.z); mc[3] = fmaxf(mc[3], c4.w);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(lo[k] > logit_cut) || !(mc[k] > logit_cut)) continue;
                const float approx = __fdividef(1.0f, (1.0f + __expf(-lo[k])) * (1.0f + __expf(-mc[k])));
                if (!(approx < screen_thr)) keep |= 1u << (4 * u + k);     // NaN class logits stay in and fail exactly
            }
        }
    }
    // one queue for the warp's survivors: exclusive prefix of the per-lane counts
    const int mine = __popc(keep);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int n_q = __shfl_sync(0xffffffffu, incl, 31);
    if (n_q == 0) return;
    {
        int pos = incl - mine;
        unsigned kb = keep;
        while (kb) {
            const int b = __ffs(kb) - 1;
            kb &= kb - 1;
            queue[pos++] = (i0 + (b >> 2) * (kScoreWarps * 32)) * 4 + (b & 3);
        }
    }
    __syncwarp();
    // exact scores, one candidate per lane
    for (int c0 = 0; c0 < n_q; c0 += 32) {
        const bool have = c0 + lane < n_q;
        bool pass = false;
        float score = 0.f;
        int label = 0, p = 0, s = -1;
        if (have) {
            const int e = queue[c0 + lane];
            s = e / hw;
            const int q = e - s * hw;
            p = p_off + q;
            const float so = st_sigmoid_f32(obj[e]);
            // scores = sigmoid(cls) * sigmoid(obj); max over classes, first max wins
            float best = -1.f;
            for (int c = 0; c < num_classes; ++c) {
                const float lc = cls[((long long)s * num_classes + c) * hw + q];
                if (!(lc > logit_cut)) continue;
                const float v = __fmul_rn(st_sigmoid_f32(lc), so);
                if (v > best) { best = v; label = c; }
            }
            score = best;
            pass = score > score_thr;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, pass);
        if (ballot == 0) continue;
        if (pass) {
            const unsigned peers = __match_any_sync(ballot, s);            // the passing lanes of the same sequence
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&cand_count[s], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            const int pos = base + __popc(peers & ((1u << lane) - 1u));
            if (pos < cap) {
                cand_score[(long long)s * cap + pos] = score;
                cand_prior[(long long)s * cap + pos] = p;
                cand_label[(long long)s * cap + pos] = label;
            }
        }
    }
}

// mmcv nms criterion: inter / (Sa + Sb - inter) > thr, offset 0.
// Boxes that do not overlap have inter = 0 and 0 / union (= +-0 or NaN) is never
// > thr for thr >= 0, so the division is only evaluated for overlapping pairs -
// nearly all of the N^2/2 pairs of a frame are disjoint.
__device__ __forceinline__ bool suppresses(const float4 a, const float4 b, float thr) {
    const float wr = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float hr = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (thr >= 0.f && (!(wr > 0.f) || !(hr > 0.f))) return false;
    const float w = fmaxf(wr, 0.f), h = fmaxf(hr, 0.f);
    const float inter = __fmul_rn(w, h);
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

__device__ __forceinline__ float clamp_like_torch(float x, float lo, float hi) {
    return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}

constexpr int kNmsThreads = 1024;

__host__ __device__ inline size_t sort_nms_smem(int cap) {
    // box (score order), 64-bit sort key, label + rank (candidate order), score / prior / label
    // (score order), removed bits, diagonal block, kept list
    return (size_t)cap * (16 + 8 + 4 + 4 + 4 + 4 + 4) + (size_t)((cap + 63) / 64) * 8 + 64 * 8 + 64 * 4 + 64;
}

__global__ void __launch_bounds__(kNmsThreads)
k2_sort_nms(HeadLevels L, const StFrameInputs *__restrict__ tab, int cap, int num_classes, float scale_w, float scale_h, float iou_thr, int det_cap,
            float ori_w, float ori_h, const int32_t *__restrict__ cand_count,
            const float *__restrict__ cand_score, const int32_t *__restrict__ cand_prior,
            const int32_t *__restrict__ cand_label, float *__restrict__ boxes, float *__restrict__ scores,
            int64_t *__restrict__ labels, int32_t *__restrict__ prior_idx, int32_t *__restrict__ keep_idx,
            int32_t *__restrict__ counts, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char k2_smem[];
    float4 *sbox = reinterpret_cast<float4 *>(k2_smem);                    // [cap] score order
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(sbox + cap);   // [cap] candidate order
    int32_t *lab = reinterpret_cast<int32_t *>(skey + cap);                // [cap] candidate order
    int32_t *rnk = lab + cap;                                              // [cap] candidate order
    float *sscore = reinterpret_cast<float *>(rnk + cap);                  // [cap] score order
    int32_t *sprior = reinterpret_cast<int32_t *>(sscore + cap);
    int32_t *slabel = sprior + cap;
    unsigned long long *remv = reinterpret_cast<unsigned long long *>(slabel + cap);   // [(cap+63)/64]
    unsigned long long *diag = remv + (cap + 63) / 64;                     // [64]
    int32_t *klist = reinterpret_cast<int32_t *>(diag + 64);               // [64] kept rows of the chunk
    __shared__ unsigned long long keepbits;
    __shared__ uint32_t max_bits;
    __shared__ int chunk_base, n_kept_chunk;

    const int s = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n = cand_count[s];
    if (n > cap) {
        if (tid == 0) atomicOr(&status[s], ST_STATUS_CAND_OVERFLOW);
        n = cap;
    }
    const long long base = (long long)s * cap;
    const int nw = (n + 63) >> 6;
    // sort key: score bits (positive floats order like their bit patterns) above the
    // inverted prior index, so that "greater key" = higher score, then lower prior index
    for (int i = tid; i < n; i += kNmsThreads) {
        skey[i] = ((unsigned long long)__float_as_uint(cand_score[base + i]) << 32) |
                  (unsigned long long)(0xffffffffu - (uint32_t)cand_prior[base + i]);
        lab[i] = cand_label[base + i];
        rnk[i] = 0;
    }
    for (int w = tid; w < nw; w += kNmsThreads) remv[w] = 0ull;
    if (tid == 0) max_bits = 0u;
    __syncthreads();
    // ---- stable rank: every thread counts the greater keys of one slice of the list ------
    if (n > 0) {
        const int parts = max(1, kNmsThreads / n);              // slices per candidate
        const int slice = (n + parts - 1) / parts;
        for (int t = tid; t < n * parts; t += kNmsThreads) {
            const int i = t % n, part = t / n;
            const unsigned long long k = skey[i];
            const int j0 = part * slice, j1 = min(n, j0 + slice);
            int cnt = 0;
#pragma unroll 4
            for (int j = j0; j < j1; ++j) cnt += skey[j] > k;
            if (cnt) atomicAdd(&rnk[i], cnt);
        }
    }
    __syncthreads();
    // ---- decode, scatter into score order -------------------------------------------------
    float mx = -3.0e38f;
    for (int i = tid; i < n; i += kNmsThreads) {
        const unsigned long long k = skey[i];
        const int p = (int)(0xffffffffu - (uint32_t)k);
        const int rank = rnk[i];
        const int l = level_of(L, p);
        const int q = p - L.prior_off[l];
        const long long hw = (long long)L.h[l] * L.w[l];
        const float *bp = K2_BBOX(l) + (long long)s * 4 * hw + q;
        const float r0 = bp[0], r1 = bp[hw], r2 = bp[2 * hw], r3 = bp[3 * hw];
        const float sf = (float)L.stride[l];
        const float px = __fmul_rn((float)(q % L.w[l]), sf);
        const float py = __fmul_rn((float)(q / L.w[l]), sf);
        // YOLOXBBoxCoder.decode: xy = p*s + prior, wh = exp(p)*s, xyxy = xy -/+ wh/2
        const float cx = __fadd_rn(__fmul_rn(r0, sf), px);
        const float cy = __fadd_rn(__fmul_rn(r1, sf), py);
        const float hwid = __fmul_rn(__fmul_rn(st_exp_f32(r2), sf), 0.5f);
        const float hhei = __fmul_rn(__fmul_rn(st_exp_f32(r3), sf), 0.5f);
        float4 b;
        b.x = __fdiv_rn(__fsub_rn(cx, hwid), scale_w);
        b.y = __fdiv_rn(__fsub_rn(cy, hhei), scale_h);
        b.z = __fdiv_rn(__fadd_rn(cx, hwid), scale_w);
        b.w = __fdiv_rn(__fadd_rn(cy, hhei), scale_h);
        sbox[rank] = b;
        sscore[rank] = __uint_as_float((uint32_t)(k >> 32));
        sprior[rank] = p;
        slabel[rank] = lab[i];
        mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    if (num_classes > 1) {
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx > -3.0e38f) atomicMax(&max_bits, ordered_bits(mx));
    }
    __syncthreads();
    // batched_nms: boxes + label * (max_coordinate + 1) keeps classes apart
    const float off_unit = num_classes > 1 ? __fadd_rn(from_ordered_bits(max_bits), 1.0f) : 0.f;
    auto nms_box = [&](int i) {
        float4 b = sbox[i];
        if (num_classes > 1) {
            const float o = __fmul_rn((float)slabel[i], off_unit);
            b = make_float4(__fadd_rn(b.x, o), __fadd_rn(b.y, o), __fadd_rn(b.z, o), __fadd_rn(b.w, o));
        }
        return b;
    };
    // ---- greedy NMS in 64-box chunks (two barriers per chunk) ---------------------------------
    int kept_total = 0;                       // thread 0 only
    for (int c = 0; c < nw; ++c) {
        const int r0 = c * 64, rows_here = min(64, n - r0);
        // (a) suppression bitmask of the chunk against itself (upper triangle), two rows per warp
        for (int r = warp; r < rows_here; r += kNmsThreads / 32) {
            const int i = r0 + r;
            const float4 a = nms_box(i);
            const int j0 = r0 + lane, j1 = j0 + 32;
            const bool p0 = j0 < n && j0 > i && suppresses(a, nms_box(j0), iou_thr);
            const bool p1 = j1 < n && j1 > i && suppresses(a, nms_box(j1), iou_thr);
            const unsigned lo = __ballot_sync(0xffffffffu, p0);
            const unsigned hi = __ballot_sync(0xffffffffu, p1);
            if (lane == 0) diag[r] = ((unsigned long long)hi << 32) | lo;
        }
        __syncthreads();          // diag complete; every thread has finished step (c) of the previous chunk
        // (b) serial resolution of the chunk: visit only the rows that are still alive
        if (tid == 0) {
            const unsigned long long live = rows_here == 64 ? ~0ull : ((1ull << rows_here) - 1ull);
            unsigned long long cur = remv[c], kb = 0ull;
            int nk = 0;
            while (true) {
                const unsigned long long todo = ~cur & live & ~kb;
                if (!todo) break;
                const int r = __ffsll((long long)todo) - 1;
                kb |= 1ull << r;
                cur |= diag[r];               // diag[r] has only bits above r
                klist[nk++] = r0 + r;
            }
            keepbits = kb;
            n_kept_chunk = nk;
            chunk_base = kept_total;
            kept_total += nk;
        }
        __syncthreads();
        const unsigned long long kb = keepbits;
        const int nk = n_kept_chunk;
        // (c) the kept boxes of the chunk against every later box
        for (int j = r0 + 64 + tid; j < n; j += kNmsThreads) {
            if ((remv[j >> 6] >> (j & 63)) & 1ull) continue;
            const float4 bj = nms_box(j);
            for (int t = 0; t < nk; ++t)
                if (suppresses(nms_box(klist[t]), bj, iou_thr)) {
                    atomicOr(&remv[j >> 6], 1ull << (j & 63));
                    break;
                }
        }
        // (d) emit the kept boxes of this chunk in order
        if (tid < 64 && ((kb >> tid) & 1ull)) {
            const int pos = chunk_base + __popcll(kb & ((1ull << tid) - 1ull));
            const int i = r0 + tid;
            if (pos < det_cap) {
                const float4 b = sbox[i];
                const long long o = (long long)s * det_cap + pos;
                // results.bboxes[:, 0::2].clamp_(0, ori_w); [:, 1::2].clamp_(0, ori_h)
                reinterpret_cast<float4 *>(boxes)[o] =
                    make_float4(clamp_like_torch(b.x, 0.f, ori_w), clamp_like_torch(b.y, 0.f, ori_h),
                                clamp_like_torch(b.z, 0.f, ori_w), clamp_like_torch(b.w, 0.f, ori_h));
                scores[o] = sscore[i];
                labels[o] = (int64_t)slabel[i];
                prior_idx[o] = sprior[i];
                keep_idx[o] = i;
            }
        }
    }
    if (tid == 0) {
        if (kept_total > det_cap) { atomicOr(&status[s], ST_STATUS_DET_OVERFLOW); kept_total = det_cap; }
        counts[s] = kept_total;
    }
}

}  // namespace

// called by st_create: opt in to the dynamic shared memory k2_sort_nms needs at this capacity
int st_k2_configure(int max_candidates) {
    const size_t need = sort_nms_smem(max_candidates);
    if (need > 220 * 1024) return ST_ERR_CAPACITY;
    return cudaFuncSetAttribute(k2_sort_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) == cudaSuccess
               ? ST_OK : ST_ERR_CUDA;
}

extern "C" int st_decode_nms(st_handle *h, const float *const *cls_host, const float *const *bbox_host,
                             const float *const *obj_host, int num_levels, const int32_t *level_h_host,
                             const int32_t *level_w_host, const int32_t *strides_host, int S,
                             int num_classes, float scale_w, float scale_h, int ori_h, int ori_w,
                             int det_cap, float *boxes, float *scores, int64_t *labels,
                             int32_t *prior_idx, int32_t *keep_idx, int32_t *counts, void *stream) {
    return st_decode_nms_ws(h, cls_host, bbox_host, obj_host, nullptr, nullptr, num_levels, level_h_host, level_w_host,
                            strides_host, S, num_classes, scale_w, scale_h, ori_h, ori_w, det_cap, boxes, scores,
                            labels, prior_idx, keep_idx, counts, stream);
}

// tab: the kernels read the head-map pointers from it (cls/bbox/obj_host then only select the
// kernel variant by their alignment); ws: candidate scratch of this frame (nullptr = the handle's)
int st_decode_nms_ws(st_handle *h, const float *const *cls_host, const float *const *bbox_host,
                     const float *const *obj_host, const StFrameInputs *tab, const StK2Scratch *ws, int num_levels,
                     const int32_t *level_h_host, const int32_t *level_w_host, const int32_t *strides_host, int S,
                     int num_classes, float scale_w, float scale_h, int ori_h, int ori_w, int det_cap, float *boxes,
                     float *scores, int64_t *labels, int32_t *prior_idx, int32_t *keep_idx, int32_t *counts,
                     void *stream) {
    if (!h || !cls_host || !bbox_host || !obj_host || num_levels < 1 || num_levels > ST_MAX_LEVELS ||
        S < 0 || S > h->cfg.max_sequences || num_classes < 1 || det_cap < 1 || !boxes || !scores ||
        !labels || !prior_idx || !keep_idx || !counts)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    StDeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    HeadLevels L;
    L.n = num_levels;
    int P = 0;
    for (int l = 0; l < num_levels; ++l) {
        L.cls[l] = cls_host[l]; L.bbox[l] = bbox_host[l]; L.obj[l] = obj_host[l];
        L.h[l] = level_h_host[l]; L.w[l] = level_w_host[l]; L.stride[l] = strides_host[l];
        L.prior_off[l] = P;
        P += level_h_host[l] * level_w_host[l];
    }
    L.prior_off[num_levels] = P;
    const int cap = h->cfg.max_candidates;
    // sigmoid(x) <= thr  <=>  x <= log(thr / (1 - thr)); a small margin keeps the cut conservative
    const double thr = (double)h->cfg.score_thr;
    const float logit_cut = (thr > 0.0 && thr < 1.0) ? (float)(log(thr / (1.0 - thr)) - 1e-3) : -3.0e38f;
    StK2Scratch own;
    own.count = h->cand_count; own.score = h->cand_score; own.prior = h->cand_prior; own.label = h->cand_label;
    if (!ws) ws = &own;
    ST_CUDA_OK(cudaMemsetAsync(ws->count, 0, sizeof(int32_t) * S, st));
    bool vec4 = true;
    for (int l = 0; l < num_levels; ++l)
        vec4 = vec4 && (level_h_host[l] * level_w_host[l]) % 4 == 0 && ((uintptr_t)cls_host[l] % 16 == 0) &&
               ((uintptr_t)obj_host[l] % 16 == 0);
    bool flat = vec4;
    for (int l = 0; l < num_levels; ++l) flat = flat && (long long)S * level_h_host[l] * level_w_host[l] < (1ll << 31);
    if (flat) {
        // 128-bit loads per thread: enough CTAs for every SM at the bench size, 32 KB per CTA in flight once
        // the pass is bandwidth-bound (2048 x 1080p, kernel alone: 85.5 / 65.0 / 53.1 us with 2 / 4 / 8)
        static const int forced = [] { const char *e = getenv("ST_K2_ITEMS"); return e ? atoi(e) : 0; }();
        const long long words = ((long long)S * P) >> 2;
        const int items = forced ? forced : (words < (1ll << 20) ? 2 : (words < (1ll << 22) ? 4 : 8));
        FlatTiles ft;
        int n_tiles = 0;
        const int per_tile = kScoreWarps * 32 * (items == 2 ? 2 : items == 8 ? 8 : 4);
        for (int l = 0; l < num_levels; ++l) {
            ft.first[l] = n_tiles;
            const long long n4 = ((long long)S * level_h_host[l] * level_w_host[l]) >> 2;
            n_tiles += (int)((n4 + per_tile - 1) / per_tile);
        }
        ft.first[num_levels] = n_tiles;
#define K2_FLAT(IT) k2_score_flat<IT><<<n_tiles, kScoreWarps * 32, 0, st>>>(L, tab, ft, S, num_classes, h->cfg.score_thr, logit_cut, cap, \
                                                                            ws->count, ws->score, ws->prior, ws->label)
        if (items == 2) K2_FLAT(2); else if (items == 8) K2_FLAT(8); else K2_FLAT(4);
#undef K2_FLAT
    } else {
        k2_score_compact<<<dim3((P + 255) / 256, S), 256, 0, st>>>(
            L, tab, P, num_classes, h->cfg.score_thr, logit_cut, cap, ws->count, ws->score, ws->prior, ws->label);
    }
    k2_sort_nms<<<S, kNmsThreads, sort_nms_smem(cap), st>>>(
        L, tab, cap, num_classes, scale_w, scale_h, h->cfg.nms_iou_thr, det_cap, (float)ori_w, (float)ori_h,
        ws->count, ws->score, ws->prior, ws->label, boxes, scores, labels, prior_idx,
        keep_idx, counts, h->st.status);
    return st_last_launch();
}
