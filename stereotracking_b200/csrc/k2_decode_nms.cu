// K2: YOLOX head post-process for S frames in lockstep:
//   k2_score_compact  prior grid, sigmoid(cls)*sigmoid(obj), max over classes,
//                     score > thr, warp-aggregated compaction
//   k2_rank_decode    stable rank by (score desc, prior index asc), box decode,
//                     rescale, scatter into score order
//   k2_nms_mask       64x64 tiles, warp-ballot suppression bitmask (upper triangle)
//   k2_nms_scan       greedy keep scan in 64-box chunks, clamp, emit
// Replaces the third-party post-process the reference calls at
// mmtrack/models/detectors/yolo_detector_disparity_v1.py:121-122 (mmyolo 0.2.0
// YOLOv5Head.predict_by_feat / YOLOXBBoxCoder.decode, mmdet 3.0.0rc4
// filter_scores_and_topk / _bbox_post_process, mmcv 2.0.0rc3 batched_nms).
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
k2_score_compact(HeadLevels L, int P, int num_classes, float score_thr, int cap,
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
        const float so = st_sigmoid_f32(L.obj[l][(long long)s * hw + q]);
        // scores = sigmoid(cls) * sigmoid(obj); max over classes, first max wins
        float best = -1.f;
        for (int c = 0; c < num_classes; ++c) {
            float sc = __fmul_rn(st_sigmoid_f32(L.cls[l][((long long)s * num_classes + c) * hw + q]), so);
            if (sc > best) { best = sc; label = c; }
        }
        score = best;
        pass = score > score_thr;
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

__global__ void __launch_bounds__(128)
k2_rank_decode(HeadLevels L, int cap, float scale_w, float scale_h,
               const int32_t *__restrict__ cand_count, const float *__restrict__ cand_score,
               const int32_t *__restrict__ cand_prior, const int32_t *__restrict__ cand_label,
               float *__restrict__ sort_box, float *__restrict__ sort_score,
               int32_t *__restrict__ sort_prior, int32_t *__restrict__ sort_label,
               uint32_t *__restrict__ max_coord, int32_t *__restrict__ status) {
    const int s = blockIdx.y;
    int n = cand_count[s];
    if (n > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&status[s], ST_STATUS_CAND_OVERFLOW);
        n = cap;
    }
    if (blockIdx.x * blockDim.x >= n) return;
    __shared__ uint32_t t_key[128];
    __shared__ int32_t t_pri[128];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const long long base = (long long)s * cap;
    uint32_t key = 0;
    int pri = 0;
    if (live) { key = __float_as_uint(cand_score[base + i]); pri = cand_prior[base + i]; }
    // scores are positive floats, so their bit patterns order like the values
    int rank = 0;
    for (int t0 = 0; t0 < n; t0 += 128) {
        int j = t0 + threadIdx.x;
        __syncthreads();
        if (j < n) { t_key[threadIdx.x] = __float_as_uint(cand_score[base + j]); t_pri[threadIdx.x] = cand_prior[base + j]; }
        __syncthreads();
        int m = min(128, n - t0);
        if (live)
            for (int k = 0; k < m; ++k)
                rank += (t_key[k] > key) || (t_key[k] == key && t_pri[k] < pri);
    }
    float mx = -3.0e38f;
    if (live) {
        const int l = level_of(L, pri);
        const int q = pri - L.prior_off[l];
        const long long hw = (long long)L.h[l] * L.w[l];
        const float *bp = L.bbox[l] + (long long)s * 4 * hw + q;
        const float sf = (float)L.stride[l];
        const float px = __fmul_rn((float)(q % L.w[l]), sf);
        const float py = __fmul_rn((float)(q / L.w[l]), sf);
        // YOLOXBBoxCoder.decode: xy = p*s + prior, wh = exp(p)*s, xyxy = xy -/+ wh/2
        const float cx = __fadd_rn(__fmul_rn(bp[0], sf), px);
        const float cy = __fadd_rn(__fmul_rn(bp[hw], sf), py);
        const float hwid = __fmul_rn(__fmul_rn(st_exp_f32(bp[2 * hw]), sf), 0.5f);
        const float hhei = __fmul_rn(__fmul_rn(st_exp_f32(bp[3 * hw]), sf), 0.5f);
        float4 b;
        b.x = __fdiv_rn(__fsub_rn(cx, hwid), scale_w);
        b.y = __fdiv_rn(__fsub_rn(cy, hhei), scale_h);
        b.z = __fdiv_rn(__fadd_rn(cx, hwid), scale_w);
        b.w = __fdiv_rn(__fadd_rn(cy, hhei), scale_h);
        reinterpret_cast<float4 *>(sort_box)[base + rank] = b;
        sort_score[base + rank] = __uint_as_float(key);
        sort_prior[base + rank] = pri;
        sort_label[base + rank] = cand_label[base + i];
        mx = fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w));
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > -3.0e38f) atomicMax(&max_coord[s], ordered_bits(mx));
}

// mmcv nms criterion: inter / (Sa + Sb - inter) > thr, offset 0
__device__ __forceinline__ bool suppresses(const float4 a, const float4 b, float thr) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

// grid (kMaskCtas, S): the CTAs of a sequence stride over the live upper-triangular
// 64x64 tiles only (a (words, words, S) grid launched 65 536 CTAs at the default
// capacity, nearly all of which found nothing to do: 45 us for ~10 live tiles).
constexpr int kMaskCtas = 16;

__global__ void __launch_bounds__(128)
k2_nms_mask(int cap, int words, float iou_thr, int num_classes,
            const int32_t *__restrict__ cand_count, const float *__restrict__ sort_box,
            const int32_t *__restrict__ sort_label, const uint32_t *__restrict__ max_coord,
            unsigned long long *__restrict__ nms_mask) {
    const int s = blockIdx.y;
    const int n = min(cand_count[s], cap);
    const int nw = (n + 63) >> 6;
    const int n_tiles = nw * (nw + 1) / 2;
    __shared__ float4 rows[64], cols[64];
    const long long base = (long long)s * cap;
    // batched_nms: boxes + label * (max_coordinate + 1) keeps classes apart
    const float off_unit = num_classes > 1 ? __fadd_rn(from_ordered_bits(max_coord[s]), 1.0f) : 0.f;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int rb = 0, rem = t;                        // tile t -> (rb, cb), cb >= rb
        while (rem >= nw - rb) { rem -= nw - rb; ++rb; }
        const int cb = rb + rem;
        __syncthreads();                            // rows / cols are reused across tiles
        if (threadIdx.x < 64) {
            int i = rb * 64 + threadIdx.x;
            if (i < n) {
                float4 b = reinterpret_cast<const float4 *>(sort_box)[base + i];
                float o = num_classes > 1 ? __fmul_rn((float)sort_label[base + i], off_unit) : 0.f;
                rows[threadIdx.x] = make_float4(__fadd_rn(b.x, o), __fadd_rn(b.y, o), __fadd_rn(b.z, o), __fadd_rn(b.w, o));
            }
        } else {
            int tt = threadIdx.x - 64, j = cb * 64 + tt;
            if (j < n) {
                float4 b = reinterpret_cast<const float4 *>(sort_box)[base + j];
                float o = num_classes > 1 ? __fmul_rn((float)sort_label[base + j], off_unit) : 0.f;
                cols[tt] = make_float4(__fadd_rn(b.x, o), __fadd_rn(b.y, o), __fadd_rn(b.z, o), __fadd_rn(b.w, o));
            }
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int j0 = cb * 64 + lane, j1 = j0 + 32;
        for (int r = warp * 16; r < warp * 16 + 16; ++r) {
            const int i = rb * 64 + r;
            if (i >= n) break;                       // uniform across the warp
            const float4 a = rows[r];
            const bool p0 = (j0 < n) && (j0 > i) && suppresses(a, cols[lane], iou_thr);
            const bool p1 = (j1 < n) && (j1 > i) && suppresses(a, cols[lane + 32], iou_thr);
            const unsigned lo = __ballot_sync(0xffffffffu, p0);
            const unsigned hi = __ballot_sync(0xffffffffu, p1);
            if (lane == 0)
                nms_mask[(base + i) * words + cb] = ((unsigned long long)hi << 32) | lo;
        }
    }
}

__device__ __forceinline__ float clamp_like_torch(float x, float lo, float hi) {
    return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}

// One CTA per sequence.  Boxes are visited in chunks of 64: the 64x64 diagonal
// block is resolved serially from shared memory, then the rows of the kept
// boxes are OR-ed into the remaining words in parallel.
__global__ void __launch_bounds__(256)
k2_nms_scan(int cap, int words, int det_cap, float ori_w, float ori_h,
            const int32_t *__restrict__ cand_count, const float *__restrict__ sort_box,
            const float *__restrict__ sort_score, const int32_t *__restrict__ sort_prior,
            const int32_t *__restrict__ sort_label, const unsigned long long *__restrict__ nms_mask,
            float *__restrict__ boxes, float *__restrict__ scores, int64_t *__restrict__ labels,
            int32_t *__restrict__ prior_idx, int32_t *__restrict__ keep_idx,
            int32_t *__restrict__ counts, int32_t *__restrict__ status) {
    extern __shared__ unsigned long long smem_u64[];
    unsigned long long *remv = smem_u64;            // [words]
    unsigned long long *diag = smem_u64 + words;    // [64]
    __shared__ unsigned long long keepbits;
    __shared__ int kept_total;
    const int s = blockIdx.x;
    const int n = min(cand_count[s], cap);
    const int nw = (n + 63) >> 6;
    const long long base = (long long)s * cap;
    for (int w = threadIdx.x; w < nw; w += blockDim.x) remv[w] = 0ull;
    if (threadIdx.x == 0) kept_total = 0;
    __syncthreads();
    for (int c = 0; c < nw; ++c) {
        const int rows_here = min(64, n - c * 64);
        if (threadIdx.x < 64)
            diag[threadIdx.x] = threadIdx.x < rows_here
                                    ? nms_mask[(base + c * 64 + threadIdx.x) * words + c] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[c], kb = 0ull;
            for (int r = 0; r < rows_here; ++r) {
                if (!((cur >> r) & 1ull)) { kb |= 1ull << r; cur |= diag[r]; }
            }
            keepbits = kb;
        }
        __syncthreads();
        const unsigned long long kb = keepbits;
        const int rest = nw - (c + 1);
        for (int idx = threadIdx.x; idx < rows_here * rest; idx += blockDim.x) {
            const int r = idx / rest, w = c + 1 + idx % rest;
            if ((kb >> r) & 1ull) {
                unsigned long long m = nms_mask[(base + c * 64 + r) * words + w];
                if (m) atomicOr(&remv[w], m);
            }
        }
        // emit the kept boxes of this chunk in order
        if (threadIdx.x < 64 && ((kb >> threadIdx.x) & 1ull)) {
            const int pos = kept_total + __popcll(kb & ((1ull << threadIdx.x) - 1ull));
            const int i = c * 64 + threadIdx.x;
            if (pos < det_cap) {
                const float4 b = reinterpret_cast<const float4 *>(sort_box)[base + i];
                const long long o = (long long)s * det_cap + pos;
                // results.bboxes[:, 0::2].clamp_(0, ori_w); [:, 1::2].clamp_(0, ori_h)
                reinterpret_cast<float4 *>(boxes)[o] =
                    make_float4(clamp_like_torch(b.x, 0.f, ori_w), clamp_like_torch(b.y, 0.f, ori_h),
                                clamp_like_torch(b.z, 0.f, ori_w), clamp_like_torch(b.w, 0.f, ori_h));
                scores[o] = sort_score[base + i];
                labels[o] = (int64_t)sort_label[base + i];
                prior_idx[o] = sort_prior[base + i];
                keep_idx[o] = i;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) kept_total += __popcll(kb);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (kept_total > det_cap) { atomicOr(&status[s], ST_STATUS_DET_OVERFLOW); kept_total = det_cap; }
        counts[s] = kept_total;
    }
}

}  // namespace

extern "C" int st_decode_nms(st_handle *h, const float *const *cls_host, const float *const *bbox_host,
                             const float *const *obj_host, int num_levels, const int32_t *level_h_host,
                             const int32_t *level_w_host, const int32_t *strides_host, int S,
                             int num_classes, float scale_w, float scale_h, int ori_h, int ori_w,
                             int det_cap, float *boxes, float *scores, int64_t *labels,
                             int32_t *prior_idx, int32_t *keep_idx, int32_t *counts, void *stream) {
    if (!h || !cls_host || !bbox_host || !obj_host || num_levels < 1 || num_levels > ST_MAX_LEVELS ||
        S < 0 || S > h->cfg.max_sequences || num_classes < 1 || det_cap < 1 || !boxes || !scores ||
        !labels || !prior_idx || !keep_idx || !counts)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
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
    const int words = (cap + 63) / 64;
    ST_CUDA_OK(cudaMemsetAsync(h->cand_count, 0, sizeof(int32_t) * S, st));
    ST_CUDA_OK(cudaMemsetAsync(h->max_coord, 0, sizeof(uint32_t) * S, st));
    k2_score_compact<<<dim3((P + 255) / 256, S), 256, 0, st>>>(
        L, P, num_classes, h->cfg.score_thr, cap, h->cand_count, h->cand_score, h->cand_prior, h->cand_label);
    k2_rank_decode<<<dim3((cap + 127) / 128, S), 128, 0, st>>>(
        L, cap, scale_w, scale_h, h->cand_count, h->cand_score, h->cand_prior, h->cand_label,
        h->sort_box, h->sort_score, h->sort_prior, h->sort_label, h->max_coord, h->st.status);
    k2_nms_mask<<<dim3(kMaskCtas, S), 128, 0, st>>>(
        cap, words, h->cfg.nms_iou_thr, num_classes, h->cand_count, h->sort_box, h->sort_label,
        h->max_coord, h->nms_mask);
    k2_nms_scan<<<S, 256, sizeof(unsigned long long) * (words + 64), st>>>(
        cap, words, det_cap, (float)ori_w, (float)ori_h, h->cand_count, h->sort_box, h->sort_score,
        h->sort_prior, h->sort_label, h->nms_mask, boxes, scores, labels, prior_idx, keep_idx,
        counts, h->st.status);
    return st_last_launch();
}
