// K1: raw u16 disparity -> depth (+ optional 3-channel disparity and mask),
// fused with nearest resize and zero pad.  One pass over HBM:
//   reads 2*H*W bytes, writes 4*Hp*Wp (+12*Hp*Wp disp3, +Hp*Wp mask) per frame.
// Replaces (reference, all Python/NumPy/torch):
//   mmtrack/datasets/transforms/loading_disparity.py:82-86,:129-134
//   mmtrack/datasets/transforms/transforms_disparity.py:52-70,:177-194
//   mmtrack/models/mot/ocsort_disparity.py:132-134
#include "st_common.cuh"

namespace {

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(uint2 *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
                 :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
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

constexpr int kUnroll = 4;    // independent 128-bit loads in flight per thread
constexpr unsigned kChunk = 256 * kUnroll;   // work items per CTA iteration

// Fast path: no resize, W and Wp multiples of 8, 16-byte aligned rows.
// A work item is 8 consecutive pixels of a padded output row (one 128-bit
// load, two 128-bit depth stores); a CTA iteration covers kChunk consecutive
// items of one image, so consecutive threads touch consecutive 16-byte words.
// All index arithmetic is 32-bit.
template <bool kDisp3, bool kMask>
__global__ void __launch_bounds__(256)
k1_disp2depth_vec8(const uint16_t *__restrict__ raw, int S, int H, int W, int Hp, int Wp,
                   float *__restrict__ depth, float *__restrict__ disp3,
                   uint8_t *__restrict__ mask, float bf, unsigned long long div_magic) {
    const unsigned w8 = W >> 3, wp8 = Wp >> 3;
    const unsigned items = (unsigned)Hp * wp8;
    const unsigned chunks_per_img = (items + kChunk - 1) / kChunk;
    const unsigned total_chunks = (unsigned)S * chunks_per_img;
    const size_t plane = (size_t)Hp * Wp;
    const float pad_depth = __fmul_rn(rcp_rn_fma(1e-6f), bf);
    for (unsigned chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const unsigned s = chunk / chunks_per_img;
        const unsigned base = (chunk - s * chunks_per_img) * kChunk + threadIdx.x;
        const uint16_t *src = raw + (size_t)s * H * W;
        float *dst = depth + (size_t)s * plane;
        {
            uint4 v[kUnroll];
            unsigned off[kUnroll];          // element offset inside the padded image
            bool live[kUnroll], inside[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const unsigned li = base + u * 256;
                // li / wp8 as a multiply-high (exact for li < 2^24, wp8 < 2^16): no I2F/MUFU
                const unsigned y = (unsigned)(((unsigned long long)li * div_magic) >> 40), x8 = li - y * wp8;
                live[u] = li < items;
                inside[u] = live[u] && y < (unsigned)H && x8 < w8;
                off[u] = li * 8;
                if (inside[u]) v[u] = ld_stream_u4(reinterpret_cast<const uint4 *>(src + (size_t)y * W + x8 * 8));
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (!live[u]) continue;
                float dz[8], dp[8];
                uint32_t mk[8];
                if (inside[u]) {
                    const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        convert(w[k] & 0xffffu, bf, dp[2 * k], dz[2 * k], mk[2 * k]);
                        convert(w[k] >> 16, bf, dp[2 * k + 1], dz[2 * k + 1], mk[2 * k + 1]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { dz[k] = pad_depth; dp[k] = 0.0f; mk[k] = 0; }
                }
                float4 *o = reinterpret_cast<float4 *>(dst + off[u]);
                st_stream_f4(o, make_float4(dz[0], dz[1], dz[2], dz[3]));
                st_stream_f4(o + 1, make_float4(dz[4], dz[5], dz[6], dz[7]));
                if (kDisp3) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float4 *q = reinterpret_cast<float4 *>(disp3 + ((size_t)s * 3 + c) * plane + off[u]);
                        st_stream_f4(q, make_float4(dp[0], dp[1], dp[2], dp[3]));
                        st_stream_f4(q + 1, make_float4(dp[4], dp[5], dp[6], dp[7]));
                    }
                }
                if (kMask) {
                    uint2 m;
                    m.x = mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
                    m.y = mk[4] | (mk[5] << 8) | (mk[6] << 16) | (mk[7] << 24);
                    st_stream_u2(reinterpret_cast<uint2 *>(mask + (size_t)s * plane + off[u]), m);
                }
            }
        }
    }
}

// General path: nearest resize (cv2 INTER_NEAREST index rule: src =
// min(floor(dst * (1/(dst_n/src_n))), src_n-1), float64) and arbitrary sizes.
__global__ void __launch_bounds__(256)
k1_disp2depth_generic(const uint16_t *__restrict__ raw, int S, int H, int W, int Hr, int Wr,
                      int Hp, int Wp, float *__restrict__ depth, float *__restrict__ disp3,
                      uint8_t *__restrict__ mask, float bf) {
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

int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

extern "C" int st_disp2depth(const uint16_t *raw, int S, int H, int W, int Hr, int Wr, int Hp,
                             int Wp, float *depth, float *disp3, uint8_t *mask, float baseline,
                             float focal_length, void *stream) {
    if (!raw || !depth || S < 0 || H <= 0 || W <= 0 || Hr <= 0 || Wr <= 0 || Hp < Hr || Wp < Wr)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float bf = (float)((double)baseline * (double)focal_length);
    const bool aligned = ((uintptr_t)raw % 16 == 0) && ((uintptr_t)depth % 16 == 0) &&
                         (!disp3 || (uintptr_t)disp3 % 16 == 0) &&
                         (!mask || (uintptr_t)mask % 8 == 0);
    const bool small = (long long)Hp * (Wp / 8) < (1ll << 24) && Wp / 8 < 65536 && (long long)H * W < (1ll << 31);
    const unsigned long long magic = ((1ull << 40) + (unsigned long long)(Wp / 8) - 1) / (unsigned long long)(Wp / 8);
    if (Hr == H && Wr == W && (W % 8 == 0) && (Wp % 8 == 0) && aligned && small) {
        const long long tiles = (long long)S * (((long long)Hp * (Wp / 8) + kChunk - 1) / kChunk);
        // grid = a multiple of the SM count: 8 resident CTAs of 256 threads per SM
        long long grid = (long long)sm_count() * 8;
        if (grid > tiles) grid = tiles;
        if (disp3 && mask)
            k1_disp2depth_vec8<true, true><<<(int)grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf, magic);
        else if (disp3)
            k1_disp2depth_vec8<true, false><<<(int)grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf, magic);
        else if (mask)
            k1_disp2depth_vec8<false, true><<<(int)grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf, magic);
        else
            k1_disp2depth_vec8<false, false><<<(int)grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf, magic);
    } else {
        const long long total = (long long)S * Hp * Wp;
        long long grid = (total + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        if (grid > cap) grid = cap;
        k1_disp2depth_generic<<<(int)grid, 256, 0, st>>>(raw, S, H, W, Hr, Wr, Hp, Wp, depth, disp3, mask, bf);
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
