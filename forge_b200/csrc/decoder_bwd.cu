// Input gradient of the fused decoder: d relu(conv_rgb(x)) / d x applied to g_rgb, in ONE fp32 kernel -- the
// pose-only backward path of reference models/volume_render.py:29-37,73 (test-time pose optimisation,
// kubric_eval.py:450-504: the decoder weights are constants, only the rendered features carry gradient).
// cuDNN runs this as three convolution backward-data calls on re-computed activations (FFT-tiled algorithms for
// the 5x5 / 6x6 filters, ~100 launches and ~40 pageable host-to-device copies per call): 1.9 ms for 5 views.
//
// The decoder is piece-wise linear, so the backward pass only needs the SIGNS of the three pre-activations.  The
// forward kernels (decoder.cu, decoder_tc.cu) emit them on request as one uint32 per output pixel
//   bits 0-15: layer-1 channels, 16-23: layer-2 channels, 24-26: rgb      (1 = positive pre-activation)
// and this kernel chains the three transposed convolutions through shared memory with the same tiling as the
// forward kernel (a CTA owns the 16x8 input pixels under a 32x16 output tile and recomputes halos):
//
//   g3 = g_rgb * [y > 0]                                   28 x 44 x 3   (zero outside the image)
//   g2 = lrelu'(m2) * corr5x5(g3, flip W3)                 24 x 40 x 8
//   g1 = lrelu'(m1) * corr5x5(g2, flip W2 s2)              20 x 36 x 16
//   gx = conv6x6 stride 2 (g1, Wt s1)                      8 x 16 x 16   -> NHWC
//
// fp32 FFMA (packed FFMA2), register-blocked 4 px x 8 channels per thread like the forward kernel; positions
// outside the image carry no gradient (they are the zero padding of the next layer, constants).
#include "common.cuh"

namespace forge {
namespace dbw {

constexpr int kThreads = 512;
constexpr int TOX = 32, TOY = 16;
constexpr int G3_H = TOY + 12, G3_W = TOX + 12, G3_PS = 4, G3_RS = G3_W * G3_PS + 4;     // 28 x 44 x (3 + pad)
constexpr int G2_H = TOY + 8, G2_W = TOX + 8, G2_PS = 12, G2_RS = G2_W * G2_PS + 4;      // 24 x 40 x 8
constexpr int G1_H = TOY + 4, G1_W = TOX + 4, G1_PS = 20, G1_RS = G1_W * G1_PS + 4;      // 20 x 36 x 16
constexpr int W3_N = 25 * 3 * 8, W2_N = 25 * 8 * 16, WD_N = 36 * 16 * 16;
constexpr int WPACK_N = W3_N + W2_N + WD_N;                                              // 13016 floats
constexpr int SM_G3 = G3_H * G3_RS, SM_G2 = G2_H * G2_RS, SM_G1 = G1_H * G1_RS;
constexpr int kSmemFloats = SM_G3 + SM_G2 + SM_G1 + WPACK_N;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bulk_load_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; ++spins) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spins > (1 << 24)) __trap();
    }
}

__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ void fma8(float2 (&a)[4], float av, const float4& w0, const float4& w1) {
    const float2 v = make_float2(av, av);
    a[0] = __ffma2_rn(v, make_float2(w0.x, w0.y), a[0]);
    a[1] = __ffma2_rn(v, make_float2(w0.z, w0.w), a[1]);
    a[2] = __ffma2_rn(v, make_float2(w1.x, w1.y), a[2]);
    a[3] = __ffma2_rn(v, make_float2(w1.z, w1.w), a[3]);
}
__device__ __forceinline__ float get8(const float2 (&a)[4], int c) { return (c & 1) ? a[c >> 1].y : a[c >> 1].x; }
// derivative of LeakyReLU(0.01) from the sign bit of the pre-activation
__device__ __forceinline__ float dlrelu(unsigned bits, int c) { return ((bits >> c) & 1u) ? 1.f : 0.01f; }

__global__ void __launch_bounds__(kThreads, 1)
decoder_bwd_data_kernel(const float* __restrict__ g_rgb, const unsigned* __restrict__ masks, const float* __restrict__ wpack,
                        float* __restrict__ g_x, int Sh, int Sw, int tiles_x) {
    extern __shared__ __align__(16) float sm[];
    float* sG3 = sm;
    float* sG2 = sG3 + SM_G3;
    float* sG1 = sG2 + SM_G2;
    float* sW3 = sG1 + SM_G1;
    float* sW2 = sW3 + W3_N;
    float* sWd = sW2 + W2_N;

    const int n = blockIdx.y;
    const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
    const int Y0 = tyi * TOY, X0 = txi * TOX;
    const int OH = 2 * Sh, OW = 2 * Sw;
    const int tid = threadIdx.x;
    const unsigned* mk = masks + static_cast<long long>(n) * OH * OW;

    __shared__ __align__(8) unsigned long long wbar;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) bulk_load_g2s(sW3, wpack, WPACK_N * sizeof(float), &wbar);

    // ---- phase A: g3 = g_rgb * [rgb pre-activation > 0] over the 28 x 44 region (zero outside the image) ----
    for (int e = tid; e < G3_H * G3_W; e += kThreads) {
        const int r = e / G3_W, c = e - r * G3_W;
        const int oy = Y0 - 6 + r, ox = X0 - 6 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) {
            const unsigned m = __ldg(mk + static_cast<long long>(oy) * OW + ox) >> 24;
            const float* gp = g_rgb + (static_cast<long long>(n) * 3 * OH + oy) * OW + ox;
            const long long cs = static_cast<long long>(OH) * OW;
            v.x = (m & 1u) ? __ldg(gp) : 0.f;
            v.y = (m & 2u) ? __ldg(gp + cs) : 0.f;
            v.z = (m & 4u) ? __ldg(gp + 2 * cs) : 0.f;
        }
        *reinterpret_cast<float4*>(sG3 + r * G3_RS + c * G3_PS) = v;
    }
    mbar_wait(&wbar, 0);
    __syncthreads();

    // ---- phase B: g2 = lrelu'(m2) * corr5x5(g3, W3 flipped), 24 x 40 x 8; task = (row, 4-px group) ----
    if (tid < G2_H * (G2_W / 4)) {
        const int row = tid % G2_H, grp = tid / G2_H;
        float2 acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int ky = 0; ky < 5; ++ky) {
            const float* grow = sG3 + (row + ky) * G3_RS + 4 * grp * G3_PS;
            float4 a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(grow + j * G3_PS);
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float* wp = sW3 + ((ky * 5 + kx) * 3 + cc) * 8;
                    const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                    for (int p = 0; p < 4; ++p) fma8(acc[p], comp(a[p + kx], cc), w0, w1);
                }
            }
        }
        const int oy = Y0 - 4 + row;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int lc = 4 * grp + p, ox = X0 - 4 + lc;
            float v[8];
            if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) {
                const unsigned m = __ldg(mk + static_cast<long long>(oy) * OW + ox) >> 16;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = get8(acc[p], c) * dlrelu(m, c);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = 0.f;
            }
            float* dst = sG2 + row * G2_RS + lc * G2_PS;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();

    // ---- phase C: g1 = lrelu'(m1) * corr5x5(g2, W2 flipped), 20 x 36 x 16; task = (row, 4-px group, ci half) ----
    if (tid < 2 * G1_H * (G1_W / 4)) {
        const int half = tid / (G1_H * (G1_W / 4)), unit = tid % (G1_H * (G1_W / 4)), row = unit % G1_H, grp = unit / G1_H;
        float2 acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int ky = 0; ky < 5; ++ky) {
            const float* grow = sG2 + (row + ky) * G2_RS + 4 * grp * G2_PS;
#pragma unroll 1
            for (int cq = 0; cq < 2; ++cq) {
                float4 a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(grow + j * G2_PS + cq * 4);
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float* wp = sW2 + ((ky * 5 + kx) * 8 + cq * 4 + cc) * 16 + half * 8;
                        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                        for (int p = 0; p < 4; ++p) fma8(acc[p], comp(a[p + kx], cc), w0, w1);
                    }
                }
            }
        }
        const int oy = Y0 - 2 + row;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int lc = 4 * grp + p, ox = X0 - 2 + lc;
            float v[8];
            if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) {
                const unsigned m = __ldg(mk + static_cast<long long>(oy) * OW + ox) >> (8 * half);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = get8(acc[p], c) * dlrelu(m, c);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = 0.f;
            }
            float* dst = sG1 + row * G1_RS + lc * G1_PS + half * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();

    // ---- phase D: gx[i] = sum_{u,v < 6} g1[2 i + (u, v)] . Wd[u][v]; task = (input row, 2-px group, ci half, u pair);
    //      the three u pairs of a unit are combined through shared memory (g3 / g2 are dead by now) ----
    {
        const int units = (TOY / 2) * (TOX / 4) * 2;                 // 8 rows x 8 pixel pairs x 2 ci halves = 128
        const bool work = tid < 3 * units;
        const int up = tid / units, unit = tid % units;
        const int iy = unit % (TOY / 2), xg = (unit / (TOY / 2)) % (TOX / 4), half = unit / ((TOY / 2) * (TOX / 4));
        float2 acc[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
        if (work) {
#pragma unroll 1
            for (int uu = 0; uu < 2; ++uu) {
                const int u = 2 * up + uu;
                const float* grow = sG1 + (2 * iy + u) * G1_RS + (4 * xg) * G1_PS;
#pragma unroll 1
                for (int cq = 0; cq < 4; ++cq) {
                    float4 a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(grow + j * G1_PS + cq * 4);
#pragma unroll
                    for (int v = 0; v < 6; ++v) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const float* wp = sWd + ((u * 6 + v) * 16 + cq * 4 + cc) * 16 + half * 8;
                            const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
                            fma8(acc[0], comp(a[v], cc), w0, w1);
                            fma8(acc[1], comp(a[v + 2], cc), w0, w1);
                        }
                    }
                }
            }
        }
        float* part = sG3 + (tid % units) * 16;                      // 2 x 128 x 16 floats <= SM_G3 + SM_G2
        if (work && up > 0) {
            float* dst = part + (up - 1) * units * 16;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                *reinterpret_cast<float4*>(dst + p * 8) = make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y);
                *reinterpret_cast<float4*>(dst + p * 8 + 4) = make_float4(acc[p][2].x, acc[p][2].y, acc[p][3].x, acc[p][3].y);
            }
        }
        __syncthreads();
        if (work && up == 0) {
            const int gy = Y0 / 2 + iy;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int gxp = X0 / 2 + 2 * xg + p;
                if (gy < Sh && gxp < Sw) {
                    const float* q1 = part + p * 8;
                    const float* q2 = part + units * 16 + p * 8;
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) o[c] = get8(acc[p], c) + q1[c] + q2[c];
                    float4* dst = reinterpret_cast<float4*>(g_x + ((static_cast<long long>(n) * Sh + gy) * Sw + gxp) * 16 + half * 8);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
    }
}

}  // namespace dbw
}  // namespace forge

extern "C" int forge_decoder_bwd_wpack_floats(void) { return forge::dbw::WPACK_N; }

extern "C" int forge_decoder_bwd_data(const float* g_rgb_nchw, const unsigned* masks, const float* wpack_bwd, float* g_x_nhwc,
                                      int N, int S_h, int S_w, void* stream) {
    FORGE_RANGE("forge_decoder_bwd_data");
    using namespace forge;
    using namespace forge::dbw;
    const char* fn = "forge_decoder_bwd_data";
    if (!g_rgb_nchw || !masks || !wpack_bwd || !g_x_nhwc) return fail(fn, "null pointer");
    if (N <= 0 || S_h <= 0 || S_w <= 0) return fail(fn, "non-positive size");
    if (N > 65535) return fail(fn, "more than 65535 images in one launch");
    if (!aligned16(wpack_bwd) || !aligned16(g_x_nhwc)) return fail(fn, "wpack_bwd / g_x_nhwc must be 16-byte aligned");
    const size_t smem = sizeof(float) * kSmemFloats;
    if (int rc = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(decoder_bwd_data_kernel), smem)) return rc;
    const int tiles_x = (2 * S_w + TOX - 1) / TOX, tiles_y = (2 * S_h + TOY - 1) / TOY;
    dim3 grid(tiles_x * tiles_y, N);
    decoder_bwd_data_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(g_rgb_nchw, masks, wpack_bwd, g_x_nhwc,
                                                                                          S_h, S_w, tiles_x);
    return check_launch(fn);
}
