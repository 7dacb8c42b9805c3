// Fused feature-image decoder: relu(conv_rgb(x)) of reference models/volume_render.py:29-37,73 in ONE
// kernel for inference (BatchNorm in eval mode, folded into the conv weights on the host):
//
//   x [N][S][S][16] (NHWC, the raymarcher's output)
//     -> ConvTranspose2d(16->16, k6, s2, p2) + BN + LeakyReLU(0.01)     [N][16][2S][2S]   (never leaves smem)
//     -> Conv2d(16->8, k5, p2)             + BN + LeakyReLU(0.01)       [N][ 8][2S][2S]   (never leaves smem)
//     -> Conv2d(8->3, k5, p2) -> ReLU                                   [N][ 3][2S][2S]   NCHW, fp32
//
// cuDNN runs this as 3 convolutions + 2 BN + 3 activations with 126 MB of intermediates through HBM
// and poorly-filled tiles (16/8/3 output channels): 1.9 ms (TF32) / 3.0 ms (fp32) at cfg-2.  Here a
// CTA owns a 32x16 output tile and recomputes the halos: input tile 22x14, layer-1 tile 40x24, layer-2
// tile 36x20, all in shared memory together with the 53 KB of weights.  fp32 FFMA on the CUDA cores
// (exact fp32, the parity bar is 1e-4 on RGB), register-blocked 4 px x 8 channels per thread so that
// every activation LDS.128 feeds 32-128 FFMA and weight reads are warp-broadcasts.
//
// The transposed conv is evaluated per output parity class: with oy = 2a + py, the taps are
// iy = a + 1 - ty, ky = py + 2 ty (ty = 0..2), same in x, i.e. four 3x3x16x16 filters.
#include "common.cuh"

namespace forge {

constexpr int kDecThreads = 512;
constexpr int TOX = 32, TOY = 16;
constexpr int IN_H = TOY / 2 + 6, IN_W = TOX / 2 + 6, IN_PS = 20, IN_RS = IN_W * IN_PS + 4;   // 14 x 22
constexpr int L1_H = TOY + 8, L1_W = TOX + 8, L1_PS = 20, L1_RS = L1_W * L1_PS + 4;           // 24 x 40
constexpr int L2_H = TOY + 4, L2_W = TOX + 4, L2_PS = 12, L2_RS = L2_W * L2_PS + 4;           // 20 x 36
constexpr int W1_N = 4 * 9 * 16 * 16, W2_N = 25 * 16 * 8, W3_N = 25 * 8 * 4;
constexpr int WPACK_N = W1_N + W2_N + W3_N + 16 + 8 + 4;                                        // 13244 floats
constexpr int SM_IN = IN_H * IN_RS, SM_L1 = L1_H * L1_RS, SM_L2 = L2_H * L2_RS;
constexpr int kDecSmemFloats = SM_IN + SM_L1 + SM_L2 + WPACK_N;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// one bulk async copy (TMA engine, SASS UBLKCP) global -> shared, completion on an mbarrier
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
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spins > (1 << 24)) __trap();     // a lost copy must fail loudly, not hang the GPU
    }
}

__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.01f * v; }

// 8 (or 4) output channels += one activation x 8 (4) weights as packed FFMA2 (two fp32 FMAs per instruction on
// sm_100; per-component rounding identical to fmaf)
__device__ __forceinline__ void fma8(float2 (&a)[4], float av, const float4& w0, const float4& w1) {
    const float2 v = make_float2(av, av);
    a[0] = __ffma2_rn(v, make_float2(w0.x, w0.y), a[0]);
    a[1] = __ffma2_rn(v, make_float2(w0.z, w0.w), a[1]);
    a[2] = __ffma2_rn(v, make_float2(w1.x, w1.y), a[2]);
    a[3] = __ffma2_rn(v, make_float2(w1.z, w1.w), a[3]);
}
__device__ __forceinline__ void fma4(float2 (&a)[2], float av, const float4& w) {
    const float2 v = make_float2(av, av);
    a[0] = __ffma2_rn(v, make_float2(w.x, w.y), a[0]);
    a[1] = __ffma2_rn(v, make_float2(w.z, w.w), a[1]);
}
__device__ __forceinline__ float get8(const float2 (&a)[4], int co) { return (co & 1) ? a[co >> 1].y : a[co >> 1].x; }

__global__ void __launch_bounds__(kDecThreads, 1)
decoder_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wpack, float* __restrict__ rgb,
                   unsigned* __restrict__ masks, int Sh, int Sw, int tiles_x) {
    extern __shared__ __align__(16) float sm[];
    float* sIn = sm;
    float* sL1 = sIn + SM_IN;
    float* sL2 = sL1 + SM_L1;
    float* sW1 = sL2 + SM_L2;
    float* sW2 = sW1 + W1_N;
    float* sW3 = sW2 + W2_N;
    float* sB1 = sW3 + W3_N;
    float* sB2 = sB1 + 16;
    float* sB3 = sB2 + 8;

    const int n = blockIdx.y;
    const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
    const int Y0 = tyi * TOY, X0 = txi * TOX;          // output-tile origin (even)
    const int OH = 2 * Sh, OW = 2 * Sw;
    const int tid = threadIdx.x;
    // sign masks for the backward pass (decoder_bwd.cu): byte 0/1 = layer-1 channels 0-7 / 8-15, byte 2 = layer 2,
    // byte 3 = rgb; every output pixel is written by the tile that owns it
    unsigned char* mbytes = masks ? reinterpret_cast<unsigned char*>(masks + static_cast<long long>(n) * OH * OW) : nullptr;

    // ---- phase 0: weights by one bulk async copy (53 KB, no register staging) overlapped with the
    //      threads' own load of the input tile (zeros outside the image) ----
    __shared__ __align__(8) unsigned long long wbar;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&wbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) bulk_load_g2s(sW1, wpack, WPACK_N * sizeof(float), &wbar);
    {
        const int iy0 = Y0 / 2 - 3, ix0 = X0 / 2 - 3;
        const float4* xin = reinterpret_cast<const float4*>(x) + static_cast<long long>(n) * Sh * Sw * 4;
        for (int e = tid; e < IN_H * IN_W * 4; e += kDecThreads) {
            const int px = e >> 2, q4 = e & 3;
            const int r = px / IN_W, c = px - r * IN_W;
            const int iy = iy0 + r, ix = ix0 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < Sh && ix >= 0 && ix < Sw) v = __ldg(xin + (static_cast<long long>(iy) * Sw + ix) * 4 + q4);
            *reinterpret_cast<float4*>(sIn + r * IN_RS + c * IN_PS + q4 * 4) = v;
        }
    }
    mbar_wait(&wbar, 0);
    __syncthreads();

    // ---- phase 1: transposed conv 16 -> 16 over the 40x24 layer-1 tile; task = (parity class, co half,
    //      4-px group, row), row fastest so that the lanes of an LDS phase hit distinct banks ----
    if (tid < 480) {
        const int r = tid % 12, g = (tid / 12) % 5, ch = tid / 60;     // ch = class * 2 + half
        const int cls = ch >> 1, half = ch & 1, py = cls >> 1, px = cls & 1;
        float2 acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < 4; ++co) acc[p][co] = make_float2(0.f, 0.f);
        const float* wbase = sW1 + cls * 2304 + half * 8;
#pragma unroll 1
        for (int ty = 0; ty < 3; ++ty) {
            const float* inrow = sIn + (r + 2 - ty) * IN_RS + 4 * g * IN_PS;
#pragma unroll 1
            for (int cq = 0; cq < 4; ++cq) {
                float4 a[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) a[j] = *reinterpret_cast<const float4*>(inrow + j * IN_PS + cq * 4);
#pragma unroll
                for (int tx = 0; tx < 3; ++tx) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float* wp = wbase + ((ty * 3 + tx) * 16 + cq * 4 + cc) * 16;
                        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                        for (int p = 0; p < 4; ++p) fma8(acc[p], comp(a[p + 2 - tx], cc), w0, w1);
                    }
                }
            }
        }
        const int lr = 2 * r + py, oy = Y0 - 4 + lr;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int lc = 2 * (4 * g + p) + px, ox = X0 - 4 + lc;
            const bool inside = (oy >= 0 && oy < OH && ox >= 0 && ox < OW);     // conv zero padding of layer 2
            float v[8];
            unsigned bits = 0;
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                const float pre = get8(acc[p], co) + sB1[half * 8 + co];
                bits |= (pre > 0.f ? 1u : 0u) << co;
                v[co] = inside ? lrelu(pre) : 0.f;
            }
            if (mbytes && inside && lr >= 4 && lr < 4 + TOY && lc >= 4 && lc < 4 + TOX)
                mbytes[(static_cast<long long>(oy) * OW + ox) * 4 + half] = static_cast<unsigned char>(bits);
            float* dst = sL1 + lr * L1_RS + lc * L1_PS + half * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();

    // ---- phase 2: conv 5x5, 16 -> 8 over the 36x20 layer-2 tile; task = (ci half, 4-px group, row); the two
    //      ci halves of a unit are combined through shared memory (sIn is dead by now) ----
    {
        const bool work = tid < 360;
        const int cih = tid / 180, unit = tid % 180, row = unit % 20, grp = unit / 20;
        float2 acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < 4; ++co) acc[p][co] = make_float2(0.f, 0.f);
        if (work) {
#pragma unroll 1
            for (int ky = 0; ky < 5; ++ky) {
                const float* l1row = sL1 + (row + ky) * L1_RS + 4 * grp * L1_PS + cih * 8;
#pragma unroll 1
                for (int cq = 0; cq < 2; ++cq) {
                    float4 a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const float4*>(l1row + j * L1_PS + cq * 4);
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const float* wp = sW2 + ((ky * 5 + kx) * 16 + cih * 8 + cq * 4 + cc) * 8;
                            const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
                            for (int p = 0; p < 4; ++p) fma8(acc[p], comp(a[p + kx], cc), w0, w1);
                        }
                    }
                }
            }
        }
        float* part = sIn + unit * 32;        // 180 * 32 floats <= SM_IN
        if (work && cih == 1) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                *reinterpret_cast<float4*>(part + p * 8) = make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y);
                *reinterpret_cast<float4*>(part + p * 8 + 4) = make_float4(acc[p][2].x, acc[p][2].y, acc[p][3].x, acc[p][3].y);
            }
        }
        __syncthreads();
        if (work && cih == 0) {
            const int oy = Y0 - 2 + row;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int lc = 4 * grp + p, ox = X0 - 2 + lc;
                const bool inside = (oy >= 0 && oy < OH && ox >= 0 && ox < OW);   // conv zero padding of layer 3
                const float4 q0 = *reinterpret_cast<const float4*>(part + p * 8), q1 = *reinterpret_cast<const float4*>(part + p * 8 + 4);
                const float o[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                float v[8];
                unsigned bits = 0;
#pragma unroll
                for (int co = 0; co < 8; ++co) {
                    const float pre = get8(acc[p], co) + o[co] + sB2[co];
                    bits |= (pre > 0.f ? 1u : 0u) << co;
                    v[co] = inside ? lrelu(pre) : 0.f;
                }
                if (mbytes && inside && row >= 2 && row < 2 + TOY && lc >= 2 && lc < 2 + TOX)
                    mbytes[(static_cast<long long>(oy) * OW + ox) * 4 + 2] = static_cast<unsigned char>(bits);
                float* dst = sL2 + row * L2_RS + lc * L2_PS;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    __syncthreads();

    // ---- phase 3: conv 5x5, 8 -> 3, ReLU, NCHW store; task = (2-px group, row) ----
    if (tid < 256) {
        const int row = tid % 16, grp = tid / 16;
        float2 acc[2][2] = {{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}};
#pragma unroll 1
        for (int ky = 0; ky < 5; ++ky) {
            const float* l2row = sL2 + (row + ky) * L2_RS + 2 * grp * L2_PS;
#pragma unroll
            for (int cq = 0; cq < 2; ++cq) {
                float4 a[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) a[j] = *reinterpret_cast<const float4*>(l2row + j * L2_PS + cq * 4);
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float4 w = *reinterpret_cast<const float4*>(sW3 + ((ky * 5 + kx) * 8 + cq * 4 + cc) * 4);
#pragma unroll
                        for (int p = 0; p < 2; ++p) fma4(acc[p], comp(a[p + kx], cc), w);
                    }
                }
            }
        }
        const int oy = Y0 + row, ox = X0 + 2 * grp;
        if (oy < OH && mbytes) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                if (ox + p < OW) {
                    const unsigned bits = (acc[p][0].x + sB3[0] > 0.f ? 1u : 0u) | (acc[p][0].y + sB3[1] > 0.f ? 2u : 0u) |
                                          (acc[p][1].x + sB3[2] > 0.f ? 4u : 0u);
                    mbytes[(static_cast<long long>(oy) * OW + ox + p) * 4 + 3] = static_cast<unsigned char>(bits);
                }
            }
        }
        if (oy < OH) {
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                float* dst = rgb + ((static_cast<long long>(n) * 3 + co) * OH + oy) * OW + ox;
                const float a0 = co == 0 ? acc[0][0].x : (co == 1 ? acc[0][0].y : acc[0][1].x);
                const float a1 = co == 0 ? acc[1][0].x : (co == 1 ? acc[1][0].y : acc[1][1].x);
                const float v0 = fmaxf(a0 + sB3[co], 0.f), v1 = fmaxf(a1 + sB3[co], 0.f);
                if (ox + 1 < OW && ((reinterpret_cast<uintptr_t>(dst) & 7u) == 0)) {
                    *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                } else {
                    if (ox < OW) dst[0] = v0;
                    if (ox + 1 < OW) dst[1] = v1;
                }
            }
        }
    }
}

}  // namespace forge

extern "C" int forge_decoder_wpack_floats(void) { return forge::WPACK_N; }

extern "C" int forge_decoder_fwd(const float* x_nhwc, const float* wpack, float* rgb_nchw, unsigned* sign_masks, int N,
                                 int S_h, int S_w, void* stream) {
    FORGE_RANGE("forge_decoder_fwd");
    using namespace forge;
    const char* fn = "forge_decoder_fwd";
    if (!x_nhwc || !wpack || !rgb_nchw) return fail(fn, "null pointer");
    if (N <= 0 || S_h <= 0 || S_w <= 0) return fail(fn, "non-positive size");
    if (N > 65535) return fail(fn, "more than 65535 images in one launch");
    if (!aligned16(x_nhwc) || !aligned16(wpack)) return fail(fn, "x_nhwc / wpack must be 16-byte aligned");
    const size_t smem = sizeof(float) * kDecSmemFloats;
    if (int rc = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(decoder_fwd_kernel), smem)) return rc;
    const int tiles_x = (2 * S_w + TOX - 1) / TOX, tiles_y = (2 * S_h + TOY - 1) / TOY;
    dim3 grid(tiles_x * tiles_y, N);
    decoder_fwd_kernel<<<grid, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(x_nhwc, wpack, rgb_nchw, sign_masks,
                                                                                       S_h, S_w, tiles_x);
    return check_launch(fn);
}
