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

// raw u16 -> (disp px, depth m, valid).  disp = raw/16 with 65535 -> 0
// (loading_disparity.py:129-134); depth = reciprocal(disp + 1e-6) * (B*f),
// torch's evaluation of `B*f / (disp + 1e-6)` (Tensor.__rtruediv__).
__device__ __forceinline__ void convert(uint32_t raw, float bf, float &disp, float &depth,
                                        uint32_t &valid) {
    valid = raw < 65535u;
    disp = valid ? __fmul_rn((float)raw, 0.0625f) : 0.0f;
    depth = __fmul_rn(__frcp_rn(__fadd_rn(disp, 1e-6f)), bf);
}

// Fast path: no resize, W and Wp multiples of 8, 16-byte aligned rows.
// One work item = 8 consecutive pixels of one padded output row.
template <bool kDisp3, bool kMask>
__global__ void __launch_bounds__(256)
k1_disp2depth_vec8(const uint16_t *__restrict__ raw, int S, int H, int W, int Hp, int Wp,
                   float *__restrict__ depth, float *__restrict__ disp3,
                   uint8_t *__restrict__ mask, float bf) {
    const int w8 = W >> 3, wp8 = Wp >> 3;
    const long long items = (long long)S * Hp * wp8;
    const long long plane = (long long)Hp * Wp;
    const float pad_depth = __fmul_rn(__frcp_rn(1e-6f), bf);
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < items;
         base += stride * U) {
        uint4 v[U];
        bool inside[U];
        long long oidx[U];
        int sidx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long it = base + u * stride;
            inside[u] = false;
            oidx[u] = -1;
            if (it < items) {
                int x8 = (int)(it % wp8);
                long long t = it / wp8;
                int y = (int)(t % Hp);
                int s = (int)(t / Hp);
                sidx[u] = s;
                oidx[u] = ((long long)s * Hp + y) * Wp + (long long)x8 * 8;
                inside[u] = (y < H) && (x8 < w8);
                if (inside[u])
                    v[u] = ld_stream_u4(reinterpret_cast<const uint4 *>(
                        raw + ((long long)s * H + y) * W + (long long)x8 * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (oidx[u] < 0) continue;
            float dz[8], dp[8];
            uint32_t mk[8];
            if (inside[u]) {
                uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    convert(w[k] & 0xffffu, bf, dp[2 * k], dz[2 * k], mk[2 * k]);
                    convert(w[k] >> 16, bf, dp[2 * k + 1], dz[2 * k + 1], mk[2 * k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) { dz[k] = pad_depth; dp[k] = 0.0f; mk[k] = 0; }
            }
            float4 *o = reinterpret_cast<float4 *>(depth + oidx[u]);
            st_stream_f4(o, make_float4(dz[0], dz[1], dz[2], dz[3]));
            st_stream_f4(o + 1, make_float4(dz[4], dz[5], dz[6], dz[7]));
            if (kDisp3) {
                long long pix = oidx[u] - (long long)sidx[u] * plane;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4 *q = reinterpret_cast<float4 *>(
                        disp3 + ((long long)sidx[u] * 3 + c) * plane + pix);
                    st_stream_f4(q, make_float4(dp[0], dp[1], dp[2], dp[3]));
                    st_stream_f4(q + 1, make_float4(dp[4], dp[5], dp[6], dp[7]));
                }
            }
            if (kMask) {
                uint2 m;
                m.x = mk[0] | (mk[1] << 8) | (mk[2] << 16) | (mk[3] << 24);
                m.y = mk[4] | (mk[5] << 8) | (mk[6] << 16) | (mk[7] << 24);
                st_stream_u2(reinterpret_cast<uint2 *>(mask + oidx[u]), m);
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

int grid_for(long long threads_needed, int block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (threads_needed + block - 1) / block;
    long long cap = (long long)sms * 8;          // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
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
    if (Hr == H && Wr == W && (W % 8 == 0) && (Wp % 8 == 0) && aligned) {
        long long items = (long long)S * Hp * (Wp / 8);
        int grid = grid_for((items + 3) / 4, 256);
        if (disp3 && mask)
            k1_disp2depth_vec8<true, true><<<grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf);
        else if (disp3)
            k1_disp2depth_vec8<true, false><<<grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf);
        else if (mask)
            k1_disp2depth_vec8<false, true><<<grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf);
        else
            k1_disp2depth_vec8<false, false><<<grid, 256, 0, st>>>(raw, S, H, W, Hp, Wp, depth, disp3, mask, bf);
    } else {
        long long total = (long long)S * Hp * Wp;
        int grid = grid_for(total, 256);
        k1_disp2depth_generic<<<grid, 256, 0, st>>>(raw, S, H, W, Hr, Wr, Hp, Wp, depth, disp3, mask, bf);
    }
    return st_last_launch();
}
