// K1: raw u16 disparity -> depth (+ optional 3-channel disparity and mask),
// fused with nearest resize and zero pad.  One pass over HBM:
//   reads 2*H*W bytes, writes 4*Hp*Wp (+12*Hp*Wp disp3, +Hp*Wp mask) per frame.
// Replaces (reference, all Python/NumPy/torch):
//   mmtrack/datasets/transforms/loading_disparity.py:82-86,:129-134
//   mmtrack/datasets/transforms/transforms_disparity.py:52-70,:177-194
//   mmtrack/models/mot/ocsort_disparity.py:132-134
#include <cstdlib>
#include "st_common.cuh"

namespace {

__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
// Correctly rounded float32 reciprocal of x = raw/16 + 1e-6 without the XU
// pipe (no MUFU): integer seed + four Newton steps in FMAs.  The input domain is
// finite (raw in [0, 65534]); tests/test_k1_reciprocal.py checks every value
// against IEEE division with exact rational arithmetic, and the GPU test feeds
// all 65536 raw codes through the kernel.
__device__ __forceinline__ float rcp_rn_fma(float x) {
    float r = __uint_as_float(0x7EF311C7u - __float_as_uint(x));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const float e = __fmaf_rn(-x, r, 1.0f);
        r = __fmaf_rn(r, e, r);
    }
    return r;
}

// raw u16 -> (disp px, depth m, valid).  disp = raw/16 with 65535 -> 0
// (loading_disparity.py:129-134); depth = reciprocal(disp + 1e-6) * (B*f),
// torch's evaluation of `B*f / (disp + 1e-6)` (Tensor.__rtruediv__).
// The u16 -> f32 conversion uses the 2^23 mantissa trick instead of I2F: the
// profile of the first version showed the XU pipe (I2F + MUFU.RCP) 94 % busy.
__device__ __forceinline__ void convert(uint32_t raw, float bf, float &disp, float &depth,
                                        uint32_t &valid) {
    valid = raw < 65535u;
    const float fr = __fsub_rn(__uint_as_float(0x4B000000u | raw), 8388608.0f);   // exact (float)raw
    disp = valid ? __fmul_rn(fr, 0.0625f) : 0.0f;
    depth = __fmul_rn(rcp_rn_fma(__fadd_rn(disp, 1e-6f)), bf);
}


// The same arithmetic on the two u16 of one 32-bit word with the packed
// float32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2 round each component
// exactly like the scalar forms): half the issue slots per pixel - the kernel is
// issue-limited, not DRAM-limited, until it is down to ~12 instructions per pixel.
// fr * 0.0625 is exact, so fma(fr, 0.0625, 1e-6) == fadd(fmul(fr, 0.0625), 1e-6).
// A pixel of raw 65535 gets disp 0 and therefore exactly the padding depth.
// (A MUFU.RCP seed + one Newton step is also bit-exact on all 65536 codes on the
// B200 and 17 % fewer instructions, but measured the same 67.7 us: after the
// packed rewrite the kernel is no longer issue-limited, so the FMA-only form,
// whose exactness tests/test_k1_reciprocal.py proves on the CPU, is kept.)
__device__ __forceinline__ void convert2(uint32_t w, float2 bf2, float pad_depth, float2 &disp, float2 &depth,
                                         uint32_t &valid_lo, uint32_t &valid_hi) {
    const uint32_t lo = __byte_perm(w, 0x4B000000u, 0x7410), hi = __byte_perm(w, 0x4B000000u, 0x7432);
    valid_lo = lo != 0x4B00FFFFu;
    valid_hi = hi != 0x4B00FFFFu;
    const float2 fr = __fadd2_rn(make_float2(__uint_as_float(lo), __uint_as_float(hi)),
                                 make_float2(-8388608.0f, -8388608.0f));
    disp = __fmul2_rn(fr, make_float2(0.0625f, 0.0625f));
    const float2 x = __ffma2_rn(fr, make_float2(0.0625f, 0.0625f), make_float2(1e-6f, 1e-6f));
    float2 r = make_float2(__uint_as_float(0x7EF311C7u - __float_as_uint(x.x)),
                           __uint_as_float(0x7EF311C7u - __float_as_uint(x.y)));
    const float2 nx = make_float2(-x.x, -x.y), one = make_float2(1.0f, 1.0f);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const float2 e = __ffma2_rn(nx, r, one);
        r = __ffma2_rn(r, e, r);
    }
    depth = __fmul2_rn(r, bf2);
    if (!valid_lo) { depth.x = pad_depth; disp.x = 0.0f; }     // 65535 -> disp 0 -> the padding depth
    if (!valid_hi) { depth.y = pad_depth; disp.y = 0.0f; }
}

// Fast path: no resize, W == Wp (no column padding), W a multiple of 32, 8-byte
// aligned images.  The padded output of all S images is one flat array of 4-pixel
// items; a warp owns 64 consecutive items per step and each lane converts item
// `lane` and item `lane + 32` of them: every warp-level access is one contiguous
// run (two 256 B loads, two 512 B stores) and the index arithmetic is shared by 8
// pixels.  (An earlier version gave each thread 8 consecutive pixels: its two
// float4 stores sat 32 B apart, every store instruction filled half of each
// 32-byte sector and the stores alone ran at 3.2 TB/s against 6.5 for a memset.)
// Because W == Wp, item li of an image reads the raw words at the same linear
// offset li while li < H * W / 4; the pad rows behave like invalid pixels (raw
// 65535 gives exactly the padding depth, disp 0 and mask 0), so they need no branch.
// The grid is the resident CTA count (SMs x occupancy) and strides over the
// flat space, so every CTA gets the same share to within one step.
// Measured on a B200 at the C2 shape (359 MB per launch, tools/bench_k1.py): 67.7 us =
// 5.3 TB/s; a torch copy of the same byte count takes 61.4 us, a memset of the
// written bytes alone 36.9 us, this kernel with its loads disabled 43 us and with
// its stores disabled 33 us.
template <bool kDisp3, bool kMask>
__global__ void __launch_bounds__(256, (kDisp3 || kMask) ? 4 : 6)
k1_disp2depth_flat(const uint16_t *__restrict__ raw, const StFrameInputs *__restrict__ tab, int S, int H, int Hp, int Wp,
                   float *__restrict__ depth, float *__restrict__ disp3,
                   uint8_t *__restrict__ mask, float bf, unsigned total_items) {
    if (tab) raw = tab->raw;                                 // replayed frame: the input lives where the table says
    constexpr int kUnroll = (kDisp3 || kMask) ? 2 : 4;      // steps in flight: 2 x 64-bit loads each
    const unsigned w4 = Wp >> 2;
    const unsigned items = (unsigned)Hp * w4, valid_items = (unsigned)H * w4;
    const unsigned stride = gridDim.x * 512u;                    // items per grid step
    const size_t plane = (size_t)Hp * Wp;
    const float pad_depth = __fmul_rn(rcp_rn_fma(1e-6f), bf);
    const float2 bf2 = make_float2(bf, bf);
    const unsigned ds = stride / items, dr = stride - ds * items;
    // this lane's first item: warp w of the CTA owns items [512 b + 64 w, +64)
    unsigned i = blockIdx.x * 512u + (threadIdx.x >> 5) * 64u + (threadIdx.x & 31u);
    unsigned s = i / items, li = i - s * items;
    const uint2 *src = reinterpret_cast<const uint2 *>(raw) + (size_t)s * valid_items + li;
    const long long src_step = (long long)ds * valid_items + dr;
    const long long src_wrap = (long long)valid_items - (long long)items;     // applied when li wraps
    while (i < total_items) {
        uint2 v[kUnroll][2];
        unsigned sq[kUnroll], lq[kUnroll];
        const unsigned i0 = i;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool live = i < total_items;
            sq[u] = s; lq[u] = li;
            v[u][0] = v[u][1] = make_uint2(0xffffffffu, 0xffffffffu);
            if (live && li < valid_items) v[u][0] = ld_stream_u2(src);
            if (live && li + 32u < valid_items) v[u][1] = ld_stream_u2(src + 32);
            i += stride;
            li += dr;
            src += src_step;
            if (li >= items) { li -= items; src += src_wrap; ++s; }
            s += ds;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const unsigned iu = i0 + u * stride;
            if (iu >= total_items) break;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                float2 dz[2], dp[2];
                uint32_t mk[4];
                convert2(v[u][hlf].x, bf2, pad_depth, dp[0], dz[0], mk[0], mk[1]);
                convert2(v[u][hlf].y, bf2, pad_depth, dp[1], dz[1], mk[2], mk[3]);
                float4 *o = reinterpret_cast<float4 *>(depth) + (size_t)iu + hlf * 32;
                st_stream_f4(o, make_float4(dz[0].x, dz[0].y, dz[1].x, dz[1].y));
                if (kDisp3) {
                    const float4 d0 = make_float4(dp[0].x, dp[0].y, dp[1].x, dp[1].y);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        reinterpret_cast<float4 *>(disp3 + ((size_t)sq[u] * 3 + c) * plane)[lq[u] + hlf * 32] = d0;
                }
                if (kMask)
                    reinterpret_cast<uint32_t *>(mask + (size_t)sq[u] * plane)[lq[u] + hlf * 32] =
                        mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
            }
        }
    }
}

// ---- resize / column-padded path: source rows staged by the bulk-copy engine ---------------
// Nearest resize (cv2 INTER_NEAREST index rule, float64: src = min(floor(dst * (1 / (dst_n /
// src_n))), src_n - 1)) and images whose width is not a multiple of 32.  An output row
// reads ONE source row, at gather positions that are not coalescible; so a CTA walks
// output rows, and one elected thread has the TMA unit copy the source row into shared
// memory (cp.async.bulk, global -> shared, completion on an mbarrier) for the row after
// the one being converted (two-stage ring).  All threads then gather from shared memory
// through a per-CTA column table and write coalesced float4 / uchar4 outputs.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K1_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K1_DONE;\n"
        "bra K1_WAIT;\n"
        "K1_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int kRowThreads = 256;

__host__ __device__ inline size_t resize_smem_bytes(int W, int Wp) {
    const size_t row = ((size_t)W * 2 + 15) & ~(size_t)15;
    return 2 * row + (((size_t)Wp * 2 + 15) & ~(size_t)15) + 16;
}

__global__ void __launch_bounds__(kRowThreads)
k1_disp2depth_rows_tma(const uint16_t *__restrict__ raw, const StFrameInputs *__restrict__ tab, int S, int H, int W,
                       int Hr, int Wr, int Hp, int Wp,
                       float *__restrict__ depth, float *__restrict__ disp3, uint8_t *__restrict__ mask, float bf) {
    if (tab) raw = tab->raw;
    extern __shared__ __align__(16) unsigned char k1_smem[];
    const size_t row_bytes = ((size_t)W * 2 + 15) & ~(size_t)15;
    uint16_t *rowbuf[2] = {reinterpret_cast<uint16_t *>(k1_smem), reinterpret_cast<uint16_t *>(k1_smem + row_bytes)};
    uint16_t *sx = reinterpret_cast<uint16_t *>(k1_smem + 2 * row_bytes);          // [Wp] source column, 0xffff = pad
    uint64_t *bars = reinterpret_cast<uint64_t *>(k1_smem + 2 * row_bytes + (((size_t)Wp * 2 + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const double inv_x = 1.0 / ((double)Wr / (double)W);
    const double inv_y = 1.0 / ((double)Hr / (double)H);
    const float pad_depth = __fmul_rn(rcp_rn_fma(1e-6f), bf);
    for (int x = tid; x < Wp; x += kRowThreads)
        sx[x] = x < Wr ? (uint16_t)min((int)floor(__dmul_rn((double)x, inv_x)), W - 1) : (uint16_t)0xffffu;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long total_rows = (long long)S * Hp;
    const size_t plane = (size_t)Hp * Wp;
    // source row of output row r (or -1 for a pad row)
    auto src_row = [&](long long r) -> long long {
        const int y = (int)(r % Hp);
        if (y >= Hr) return -1;
        const int sy = min((int)floor(__dmul_rn((double)y, inv_y)), H - 1);
        return (r / Hp) * H + sy;
    };
    unsigned phase[2] = {0u, 0u};
    long long r = blockIdx.x;
    if (r < total_rows && tid == 0) {
        const long long sr = src_row(r);
        if (sr >= 0) {
            mbar_expect_tx(&bars[0], (unsigned)row_bytes);
            bulk_g2s(rowbuf[0], raw + sr * W, (unsigned)row_bytes, &bars[0]);
        }
    }
    int stage = 0;
    for (; r < total_rows; r += gridDim.x, stage ^= 1) {
        // prefetch the next row of this CTA into the other stage (its previous reader finished
        // before the __syncthreads at the end of the previous iteration)
        const long long rn = r + gridDim.x;
        if (rn < total_rows && tid == 0) {
            const long long srn = src_row(rn);
            if (srn >= 0) {
                mbar_expect_tx(&bars[stage ^ 1], (unsigned)row_bytes);
                bulk_g2s(rowbuf[stage ^ 1], raw + srn * W, (unsigned)row_bytes, &bars[stage ^ 1]);
            }
        }
        const bool has_src = src_row(r) >= 0;
        if (has_src) {
            mbar_wait(&bars[stage], phase[stage]);
            phase[stage] ^= 1u;
        }
        const uint16_t *row = rowbuf[stage];
        const int s = (int)(r / Hp);
        const size_t out0 = (size_t)r * Wp;                     // == (s * Hp + y) * Wp
        const size_t pix0 = out0 - (size_t)s * plane;            // offset inside the image plane
        for (int x = 4 * tid; x < Wp; x += 4 * kRowThreads) {
            float dz[4], dp[4];
            uint32_t mk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = has_src ? sx[x + k] : 0xffffu;
                if (c != 0xffffu) convert(row[c], bf, dp[k], dz[k], mk[k]);
                else { dz[k] = pad_depth; dp[k] = 0.0f; mk[k] = 0u; }
            }
            st_stream_f4(reinterpret_cast<float4 *>(depth + out0 + x), make_float4(dz[0], dz[1], dz[2], dz[3]));
            if (disp3) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    *reinterpret_cast<float4 *>(disp3 + ((size_t)s * 3 + c) * plane + pix0 + x) =
                        make_float4(dp[0], dp[1], dp[2], dp[3]);
            }
            if (mask)
                *reinterpret_cast<uint32_t *>(mask + out0 + x) = mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
        }
        __syncthreads();            // every thread is done with rowbuf[stage] before it is refilled
    }
}

// General path: nearest resize (cv2 INTER_NEAREST index rule: src =
// min(floor(dst * (1/(dst_n/src_n))), src_n-1), float64) and arbitrary sizes.
__global__ void __launch_bounds__(256)
k1_disp2depth_generic(const uint16_t *__restrict__ raw, const StFrameInputs *__restrict__ tab, int S, int H, int W,
                      int Hr, int Wr,
                      int Hp, int Wp, float *__restrict__ depth, float *__restrict__ disp3,
                      uint8_t *__restrict__ mask, float bf) {
    if (tab) raw = tab->raw;
    const long long total = (long long)S * Hp * Wp;
    const long long plane = (long long)Hp * Wp;
    const double inv_x = 1.0 / ((double)Wr / (double)W);
    const double inv_y = 1.0 / ((double)Hr / (double)H);
    const float pad_depth = __fmul_rn(__frcp_rn(1e-6f), bf);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % Wp);
        long long t = i / Wp;
        int y = (int)(t % Hp);
        int s = (int)(t / Hp);
        float dp = 0.0f, dz = pad_depth;
        uint32_t mk = 0;
        if (y < Hr && x < Wr) {
            int sx = min((int)floor(__dmul_rn((double)x, inv_x)), W - 1);
            int sy = min((int)floor(__dmul_rn((double)y, inv_y)), H - 1);
            convert(raw[((long long)s * H + sy) * W + sx], bf, dp, dz, mk);
        }
        depth[i] = dz;
        if (disp3) {
            long long pix = i - (long long)s * plane;
            for (int c = 0; c < 3; ++c) disp3[((long long)s * 3 + c) * plane + pix] = dp;
        }
        if (mask) mask[i] = (uint8_t)mk;
    }
}

// Depth from an already post-processed float disparity (the `disp_postp` tensor the
// reference model receives, ocsort_disparity.py:115,:132-134).  Arbitrary floats, so
// the IEEE reciprocal is used here (not the finite-domain FMA sequence above).
__global__ void __launch_bounds__(256)
k1_disp_f32_to_depth(const float *__restrict__ disp, long long n, float *__restrict__ depth, float bf) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        depth[i] = __fmul_rn(__frcp_rn(__fadd_rn(disp[i], 1e-6f)), bf);
}

int sm_count() { return st_ctx_sms(); }

}  // namespace

extern "C" int st_disp2depth(const uint16_t *raw, int S, int H, int W, int Hr, int Wr, int Hp,
                             int Wp, float *depth, float *disp3, uint8_t *mask, float baseline,
                             float focal_length, void *stream) {
    return st_disp2depth_tab(raw, nullptr, S, H, W, Hr, Wr, Hp, Wp, depth, disp3, mask, baseline, focal_length, stream);
}

// `raw` selects the kernel variant (alignment); with a table the kernels read tab->raw instead,
// which must have the same alignment class
int st_disp2depth_tab(const uint16_t *raw, const StFrameInputs *tab, int S, int H, int W, int Hr, int Wr, int Hp,
                      int Wp, float *depth, float *disp3, uint8_t *mask, float baseline,
                      float focal_length, void *stream) {
    if (!raw || !depth || S < 0 || H <= 0 || W <= 0 || Hr <= 0 || Wr <= 0 || Hp < Hr || Wp < Wr)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float bf = (float)((double)baseline * (double)focal_length);
    const bool aligned = ((uintptr_t)raw % 8 == 0) && ((uintptr_t)depth % 16 == 0) &&
                         (!disp3 || (uintptr_t)disp3 % 16 == 0) && (!mask || (uintptr_t)mask % 4 == 0);
    const long long total_items = (long long)S * Hp * (Wp / 4);
    const bool small = total_items < (1ll << 31) - (1ll << 26);
    if (Hr == H && Wr == W && W == Wp && (Wp % 4 == 0) && (((long long)Hp * (Wp / 4)) % 64 == 0) && aligned && small) {
        // grid = the resident CTA count (SMs x occupancy of this instantiation), never more
        auto launch = [&](auto kernel) {
            struct Dev { long long resident; };
            static StPerDevice<Dev> per_device;          // per instantiation and per CUDA device
            const Dev *dv = per_device.get([&](Dev &v) {
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1)
                    per_sm = 4;
                if (const char *e = getenv("ST_K1_CTAS_PER_SM")) { const int c = atoi(e); if (c > 0 && c < per_sm) per_sm = c; }
                v.resident = (long long)st_ctx_sms() * per_sm;
                return true;
            });
            long long grid = dv ? dv->resident : 148;
            const long long need = (total_items + 511) / 512;
            if (grid > need) grid = need;
            kernel<<<(int)grid, 256, 0, st>>>(raw, tab, S, H, Hp, Wp, depth, disp3, mask, bf, (unsigned)total_items);
        };
        if (disp3 && mask) launch(k1_disp2depth_flat<true, true>);
        else if (disp3) launch(k1_disp2depth_flat<true, false>);
        else if (mask) launch(k1_disp2depth_flat<false, true>);
        else launch(k1_disp2depth_flat<false, false>);
    } else if ((W % 8 == 0) && (Wp % 4 == 0) && W <= 32768 && Wp <= 65534 && ((uintptr_t)raw % 16 == 0) &&
               ((uintptr_t)depth % 16 == 0) && (!disp3 || (uintptr_t)disp3 % 16 == 0) &&
               (!mask || (uintptr_t)mask % 4 == 0)) {
        // resize and / or column padding: TMA-staged source rows
        const size_t smem = resize_smem_bytes(W, Wp);
        // the opt-in is per device and cheap: set it for the size this call needs
        ST_CUDA_OK(cudaFuncSetAttribute(k1_disp2depth_rows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_disp2depth_rows_tma, kRowThreads, smem) != cudaSuccess ||
            per_sm < 1)
            per_sm = 1;
        long long grid = (long long)sm_count() * per_sm;
        const long long rows = (long long)S * Hp;
        if (grid > rows) grid = rows;
        k1_disp2depth_rows_tma<<<(int)grid, kRowThreads, smem, st>>>(raw, tab, S, H, W, Hr, Wr, Hp, Wp, depth, disp3, mask, bf);
    } else {
        const long long total = (long long)S * Hp * Wp;
        long long grid = (total + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        if (grid > cap) grid = cap;
        k1_disp2depth_generic<<<(int)grid, 256, 0, st>>>(raw, tab, S, H, W, Hr, Wr, Hp, Wp, depth, disp3, mask, bf);
    }
    return st_last_launch();
}

extern "C" int st_disp_f32_to_depth(const float *disp, long long n, float *depth, float baseline,
                                    float focal_length, void *stream) {
    if (!disp || !depth || n < 0) return ST_ERR_INVALID_ARG;
    if (n == 0) return ST_OK;
    long long grid = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    k1_disp_f32_to_depth<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(
        disp, n, depth, (float)((double)baseline * (double)focal_length));
    return st_last_launch();
}
