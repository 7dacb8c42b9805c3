// 3x3x3 convolution (stride 1, zero padding 1) over channels-last volumes as an implicit GEMM on the tcgen05 tensor cores,
// with the ConvGRU gate arithmetic fused into the epilogue.
//
// Replaces the cuDNN convolutions + elementwise chain of reference models/fusion.py:18-35 (ConvGRUCell_3D: conv_gate 256 ->
// 256, out_gate 256 -> 128, sigmoid / split / mul / cat / tanh / lerp) and :61-68 (fusion_conv: two conv + BN + LeakyReLU)
// on the inference / test-time-refinement path: bf16 operands, fp32 accumulation in TMEM, fp32 recurrent state.
//
//   GEMM view    M = voxels (tile = 4 x 4 x 8 block of z, y, x = 128 rows), N = Cout (128 or 256, one tile),
//                K = 27 taps x Cin, walked in steps of 64 channels (= one 128-byte swizzle row)
//   A operand    activations [n][D][H][W][C] bf16; the tile of tap (dz, dy, dx) is the SAME 5-D TMA box shifted by the tap
//                -- out-of-volume coordinates are zero-filled by the TMA unit, which is the convolution's zero padding.
//                No im2col, no halo handling in the kernel.  Two source tensors (x_t and h) are walked back to back, so
//                torch.cat([x, h]) (fusion.py:29, :33) is never materialised.
//   B operand    weights prepacked [tap][Cin / 64][Cout][64] bf16 (K-major rows of 128 bytes), one 3-D TMA box per K step
//   pipeline     warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue; kStages-deep
//                full / empty mbarrier ring for the operands, two TMEM accumulators (2 x BN columns) so the epilogue of
//                tile i overlaps the MMAs of tile i + 1; persistent CTAs, one per SM.
//   epilogues    PLAIN  y = act(acc * scale + shift)                      (conv + folded BN + LeakyReLU; fusion_conv)
//                GATE   u = sigmoid(acc[:, :C] + b), r = sigmoid(acc[:, C:] + b): stores u (fp32) and h * r (bf16, the second
//                       source of the out-gate convolution)
//                OUT    c = tanh(acc + b); h' = h (1 - u) + c u: stores h' fp32 (state) + bf16 (next step's operand), and,
//                       on the last step, fusion_norm(h') (BN eval) as the module output
#include <cuda_bf16.h>

#include <cstdlib>

#include "tensormap.cuh"

namespace forge {
namespace c3d {

using namespace async_;

constexpr int BM = 128, BK = 64;                   // voxels per tile, channels per K step
constexpr int TZ = 4, TY = 4, TX = 8;              // the tile's voxel block
constexpr int kThreads = 6 * 32;
constexpr int A_BYTES = BM * BK * 2;               // 16 KB

template <int BN>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM = STAGES * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
};

struct Params {
    int B, D, H, W, Cout;
    int nchunk_x, nchunk;          // 64-channel chunks of the first source / of both sources
    int mode;                      // 0 PLAIN, 1 GATE, 2 OUT, 3 SHUFFLE (x2 pixel shuffle of 8 x 32 columns), 4 HEADS (split outputs)
    int out_pitch, out_off;        // SHUFFLE: channels per output voxel row / first channel written; HEADS: split column
    int lrelu;                     // PLAIN: apply LeakyReLU(0.01)
    const float* scale;            // PLAIN: per-channel scale (nullable = 1); OUT: fusion_norm scale (nullable = no norm output)
    const float* shift;            // PLAIN: per-channel shift (bias folded in); GATE / OUT: conv bias; OUT: see norm_shift
    const float* norm_shift;       // OUT: fusion_norm shift
    const float* h;                // GATE / OUT: fp32 state [B][D][H][W][C]
    const float* u_in;             // OUT: update gate from the GATE launch
    float* out_f32;                // PLAIN: y fp32 (nullable); GATE: u; OUT: h'
    __nv_bfloat16* out_bf16;       // PLAIN: y bf16 (nullable); GATE: h * r; OUT: h' bf16 (nullable)
    float* out_norm;               // OUT: fusion_norm(h') (nullable)
    float* aux;                    // GATE: reset gate r; OUT: candidate c = tanh(.) (nullable; saved for the backward pass)
};

__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    // K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO), LBO unused (1),
    // descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct Maps {
    CUtensorMap x, h, w;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
    using C = Cfg<BN>;
    extern __shared__ unsigned char smem_raw[];
    // operand tiles must sit on 1024-byte boundaries (the 128-byte swizzle atom is 8 rows x 128 bytes)
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* aligned = smem_raw + (base - smem_u32(smem_raw));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(aligned + C::STAGES * C::STAGE);
    unsigned long long* full = bars;                          // [STAGES]
    unsigned long long* empty = bars + C::STAGES;             // [STAGES]
    unsigned long long* tfull = bars + 2 * C::STAGES;         // [2]  accumulator ready
    unsigned long long* tempty = tfull + 2;                   // [2]  accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = p.W / TX, tiles_y = p.H / TY, tiles_z = p.D / TZ;
    const int tiles_per_vol = tiles_x * tiles_y * tiles_z, total_tiles = tiles_per_vol * p.B;
    const int ksteps = 27 * p.nchunk;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        mbar_init_fence();
    }
    if (warp == 1) {            // TMEM: two accumulators of BN fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            prefetch_tensormap(&maps.x);
            prefetch_tensormap(&maps.h);
            prefetch_tensormap(&maps.w);
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_vol, t2 = tile - n * tiles_per_vol;
                const int tz = t2 / (tiles_x * tiles_y), t3 = t2 - tz * tiles_x * tiles_y;
                const int tyi = t3 / tiles_x, txi = t3 - tyi * tiles_x;
                const int z0 = tz * TZ, y0 = tyi * TY, x0 = txi * TX;
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const int st = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    mbar_wait(&empty[st], ph ^ 1);
                    const int tap = ks / p.nchunk, cc = ks - tap * p.nchunk;
                    const int dz = tap / 9, dy = (tap - dz * 9) / 3, dx = tap - dz * 9 - dy * 3;
                    const uint32_t sa = base + st * C::STAGE, sb = sa + A_BYTES;
                    mbar_arrive_expect_tx(&full[st], C::STAGE);
                    const bool from_x = cc < p.nchunk_x;
                    tma_load_5d(sa, from_x ? &maps.x : &maps.h, (from_x ? cc : cc - p.nchunk_x) * BK, x0 + dx - 1, y0 + dy - 1,
                                z0 + dz - 1, n, &full[st]);
                    tma_load_3d(sb, &maps.w, 0, 0, ks, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // instruction descriptor (kind::f16): D = f32, A = B = bf16, both K-major, N = BN, M = 128
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(BN) >> 3) << 17) | ((128u >> 4) << 24);
        int it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem + acc * BN;
            for (int ks = 0; ks < ksteps; ++ks, ++it) {
                const int st = it % C::STAGES;
                mbar_wait(&full[st], (it / C::STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = base + st * C::STAGE, sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, sw128_desc(sa + k * 32), sw128_desc(sb + k * 32), idesc, (ks | k) != 0);
                    umma_commit(&empty[st]);                           // frees the stage when these MMAs have read it
                    if (ks == ksteps - 1) umma_commit(&tfull[acc]);    // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue warps (TMEM lane quarter = warp % 4) =================
        const int qd = warp & 3;
        const int row = qd * 32 + lane;                                // accumulator row = voxel of the tile
        const int Cg = (p.mode == 1) ? p.Cout / 2 : p.Cout;             // channels of the state / output tensors
        int tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int acc = tcount & 1;
            const int n = tile / tiles_per_vol, t2 = tile - n * tiles_per_vol;
            const int tz = t2 / (tiles_x * tiles_y), t3 = t2 - tz * tiles_x * tiles_y;
            const int tyi = t3 / tiles_x, txi = t3 - tyi * tiles_x;
            const int z = tz * TZ + (row >> 5), y = tyi * TY + ((row >> 3) & 3), x = txi * TX + (row & 7);
            const long long vox = ((static_cast<long long>(n) * p.D + z) * p.H + y) * p.W + x;
            mbar_wait(&tfull[acc], (tcount >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            const uint32_t t_row = tmem + (static_cast<uint32_t>(qd * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tmem_ld32(t_row + c0, v);
                if (p.mode == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        float t = v[e] * (p.scale ? __ldg(p.scale + c0 + e) : 1.f) + __ldg(p.shift + c0 + e);
                        v[e] = p.lrelu ? fmaxf(t, 0.01f * t) : t;
                    }
                    if (p.out_f32) {
                        float4* o = reinterpret_cast<float4*>(p.out_f32 + vox * Cg + c0);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    }
                    if (p.out_bf16) {
                        uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + vox * Cg + c0);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            o[e] = make_uint4(pack_bf16(v[8 * e], v[8 * e + 1]), pack_bf16(v[8 * e + 2], v[8 * e + 3]),
                                              pack_bf16(v[8 * e + 4], v[8 * e + 5]), pack_bf16(v[8 * e + 6], v[8 * e + 7]));
                    }
                } else if (p.mode == 1) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = sigmoidf_(v[e] + __ldg(p.shift + c0 + e));
                    if (c0 < Cg) {              // update gate u -> fp32
                        float4* o = reinterpret_cast<float4*>(p.out_f32 + vox * Cg + c0);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    } else {                    // reset gate r -> h * r as the bf16 operand of the out-gate convolution
                        const float4* hp = reinterpret_cast<const float4*>(p.h + vox * Cg + (c0 - Cg));
                        uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + vox * Cg + (c0 - Cg));
                        if (p.aux) {
                            float4* ra = reinterpret_cast<float4*>(p.aux + vox * Cg + (c0 - Cg));
#pragma unroll
                            for (int e = 0; e < 8; ++e) ra[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 h0 = __ldg(hp + 2 * e), h1 = __ldg(hp + 2 * e + 1);
                            o[e] = make_uint4(pack_bf16(h0.x * v[8 * e], h0.y * v[8 * e + 1]), pack_bf16(h0.z * v[8 * e + 2], h0.w * v[8 * e + 3]),
                                              pack_bf16(h1.x * v[8 * e + 4], h1.y * v[8 * e + 5]),
                                              pack_bf16(h1.z * v[8 * e + 6], h1.w * v[8 * e + 7]));
                        }
                    }
                } else if (p.mode == 3) {
                    // transposed convolution (k 4, stride 2, pad 1) as a 27-tap GEMM whose 8 column groups of 32 are the 8
                    // output parity classes: columns [32 q, 32 q + 32) of input voxel (z, y, x) are the 32 channels of output
                    // voxel (2z + qz, 2y + qy, 2x + qx); BN + LeakyReLU folded, stored bf16 at channel offset out_off of rows
                    // of out_pitch channels
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float t = v[e] * __ldg(p.scale + c0 + e) + __ldg(p.shift + c0 + e);
                        v[e] = p.lrelu ? fmaxf(t, 0.01f * t) : t;
                    }
                    const int q = c0 >> 5;
                    const long long ovox = ((static_cast<long long>(n) * (2 * p.D) + (2 * z + (q >> 2))) * (2 * p.H) + (2 * y + ((q >> 1) & 1))) *
                                               (2 * p.W) + (2 * x + (q & 1));
                    uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + ovox * p.out_pitch + p.out_off);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        o[e] = make_uint4(pack_bf16(v[8 * e], v[8 * e + 1]), pack_bf16(v[8 * e + 2], v[8 * e + 3]),
                                          pack_bf16(v[8 * e + 4], v[8 * e + 5]), pack_bf16(v[8 * e + 6], v[8 * e + 7]));
                } else if (p.mode == 4) {
                    // two heads in one 32-column GEMM: columns [0, 16) -> out_f32 rows of 16 (affine, no activation: the render
                    // features), columns [16, 24) -> aux rows of 8 (affine + LeakyReLU: the density branch)
#pragma unroll
                    for (int e = 0; e < 24; ++e) v[e] = v[e] * __ldg(p.scale + e) + __ldg(p.shift + e);
                    float4* o = reinterpret_cast<float4*>(p.out_f32 + vox * 16);
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    float4* a = reinterpret_cast<float4*>(p.aux + vox * 8);
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        a[e] = make_float4(fmaxf(v[16 + 4 * e], 0.01f * v[16 + 4 * e]), fmaxf(v[17 + 4 * e], 0.01f * v[17 + 4 * e]),
                                           fmaxf(v[18 + 4 * e], 0.01f * v[18 + 4 * e]), fmaxf(v[19 + 4 * e], 0.01f * v[19 + 4 * e]));
                } else {
                    const float4* hp = reinterpret_cast<const float4*>(p.h + vox * Cg + c0);
                    const float4* up = reinterpret_cast<const float4*>(p.u_in + vox * Cg + c0);
                    float4* ca = p.aux ? reinterpret_cast<float4*>(p.aux + vox * Cg + c0) : nullptr;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 h4 = __ldg(hp + e), u4 = __ldg(up + e);
                        const float hh[4] = {h4.x, h4.y, h4.z, h4.w}, uu[4] = {u4.x, u4.y, u4.z, u4.w};
                        float cc[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            cc[t] = tanhf(v[4 * e + t] + __ldg(p.shift + c0 + 4 * e + t));
                            v[4 * e + t] = hh[t] * (1.f - uu[t]) + cc[t] * uu[t];        // reference fusion.py:35
                        }
                        if (ca) ca[e] = make_float4(cc[0], cc[1], cc[2], cc[3]);
                    }
                    float4* o = reinterpret_cast<float4*>(p.out_f32 + vox * Cg + c0);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    if (p.out_bf16) {
                        uint4* ob = reinterpret_cast<uint4*>(p.out_bf16 + vox * Cg + c0);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            ob[e] = make_uint4(pack_bf16(v[8 * e], v[8 * e + 1]), pack_bf16(v[8 * e + 2], v[8 * e + 3]),
                                               pack_bf16(v[8 * e + 4], v[8 * e + 5]), pack_bf16(v[8 * e + 6], v[8 * e + 7]));
                    }
                    if (p.out_norm) {
                        float4* on = reinterpret_cast<float4*>(p.out_norm + vox * Cg + c0);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float w4[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                w4[t] = v[4 * e + t] * __ldg(p.scale + c0 + 4 * e + t) + __ldg(p.norm_shift + c0 + 4 * e + t);
                            on[e] = make_float4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN) : "memory");
}

// Conv3d(8 -> 1, 3x3x3, pad 1) + ReLU over channels-last rows of 8 floats: the last layer of the density head (reference
// models/encoder.py:32-33).  0.45 GFLOP per 64^3 object: one thread per output voxel, 27 x 2 float4 neighbour loads (L1 / L2 hits).
__global__ void __launch_bounds__(256) conv_c8_to_1_relu_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                                 float bias, float* __restrict__ y, int B, int D, int H, int W) {
    __shared__ float sw[27 * 8];
    for (int e = threadIdx.x; e < 27 * 8; e += 256) sw[e] = w[e];
    __syncthreads();
    const long long total = static_cast<long long>(B) * D * H * W;
    for (long long o = blockIdx.x * 256ll + threadIdx.x; o < total; o += static_cast<long long>(gridDim.x) * 256) {
        const int xw = static_cast<int>(o % W), yh = static_cast<int>((o / W) % H), zd = static_cast<int>((o / (static_cast<long long>(W) * H)) % D);
        const long long nb = o / (static_cast<long long>(W) * H * D);
        float acc = bias;
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const int zz = zd + t / 9 - 1, yy = yh + (t / 3) % 3 - 1, xx = xw + t % 3 - 1;
            if (zz < 0 || zz >= D || yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float4* px = x + (((nb * D + zz) * H + yy) * W + xx) * 2;
            const float4 a = __ldg(px), b = __ldg(px + 1);
            const float* wt = sw + t * 8;
            acc = fmaf(a.x, wt[0], fmaf(a.y, wt[1], fmaf(a.z, wt[2], fmaf(a.w, wt[3], acc))));
            acc = fmaf(b.x, wt[4], fmaf(b.y, wt[5], fmaf(b.z, wt[6], fmaf(b.w, wt[7], acc))));
        }
        y[o] = fmaxf(acc, 0.f);
    }
}

static int act_map(const char* fn, CUtensorMap* m, const void* ptr, long long batch_stride, int C, int B, int D, int H, int W) {
    const unsigned long long dims[5] = {static_cast<unsigned long long>(C), static_cast<unsigned long long>(W),
                                        static_cast<unsigned long long>(H), static_cast<unsigned long long>(D),
                                        static_cast<unsigned long long>(B)};
    const unsigned long long strides[4] = {2ull * C, 2ull * C * W, 2ull * C * W * H, 2ull * static_cast<unsigned long long>(batch_stride)};
    const unsigned box[5] = {BK, TX, TY, TZ, 1};
    return encode_tensor_map(fn, m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace c3d
}  // namespace forge

extern "C" int forge_conv3d_tc(const void* x, long long x_batch_stride, int Cx, const void* h2, long long h_batch_stride,
                               int Ch, const void* wpack, int mode, int lrelu, const float* scale, const float* shift,
                               const float* norm_shift, const float* h_state, const float* u_in, float* out_f32,
                               void* out_bf16, float* out_norm, float* aux, int out_pitch, int out_off, int B, int D, int H, int W,
                               int Cout, int max_ctas, void* stream) {
    FORGE_RANGE("forge_conv3d_tc");
    using namespace forge;
    using namespace forge::c3d;
    const char* fn = "forge_conv3d_tc";
    if (!x || !wpack || !shift) return fail(fn, "null pointer");
    if (B <= 0 || D % TZ || H % TY || W % TX || D <= 0 || H <= 0 || W <= 0) return fail(fn, "volume sides must be multiples of 4 (z, y) and 8 (x)");
    if (Cx <= 0 || Cx % BK || Ch < 0 || Ch % BK || (Ch > 0 && !h2)) return fail(fn, "source channels must be multiples of 64");
    if (Cout != 32 && Cout != 128 && Cout != 256) return fail(fn, "Cout must be 32, 128 or 256");
    if (mode < 0 || mode > 4) return fail(fn, "mode must be 0 (plain), 1 (gate), 2 (out), 3 (shuffle) or 4 (heads)");
    if (mode == 3 && (Cout != 256 || !out_bf16 || !scale || out_pitch < 32 || out_off < 0 || out_off + 32 > out_pitch || (out_pitch | out_off) % 8))
        return fail(fn, "shuffle mode needs Cout = 256 (8 classes x 32), scale, a bf16 output and 8-aligned out_pitch / out_off");
    if (mode == 4 && (Cout != 32 || !out_f32 || !aux || !scale)) return fail(fn, "heads mode needs Cout = 32, scale, out_f32 [..][16] and aux [..][8]");
    if (mode != 4 && mode != 0 && Cout == 32) return fail(fn, "Cout = 32 is supported in plain and heads mode only");
    if (mode == 0 && !out_f32 && !out_bf16) return fail(fn, "plain mode needs an output");
    if (mode == 1 && (!h_state || !out_f32 || !out_bf16)) return fail(fn, "gate mode needs h, u (fp32) and h*r (bf16) buffers");
    if (mode == 2 && (!h_state || !u_in || !out_f32)) return fail(fn, "out mode needs h, u and the new-state buffer");
    if (mode == 2 && out_norm && (!scale || !norm_shift)) return fail(fn, "out mode with a norm output needs scale and norm_shift");
    if (!aligned16(x) || (h2 && !aligned16(h2)) || !aligned16(wpack)) return fail(fn, "operands must be 16-byte aligned");
    Maps maps;
    if (int e = act_map(fn, &maps.x, x, x_batch_stride, Cx, B, D, H, W)) return e;
    if (int e = act_map(fn, &maps.h, Ch ? h2 : x, Ch ? h_batch_stride : x_batch_stride, Ch ? Ch : Cx, B, D, H, W)) return e;
    Params p;
    p.B = B, p.D = D, p.H = H, p.W = W, p.Cout = Cout;
    p.nchunk_x = Cx / BK, p.nchunk = (Cx + Ch) / BK;
    p.mode = mode, p.lrelu = lrelu;
    p.out_pitch = out_pitch, p.out_off = out_off;
    p.scale = scale, p.shift = shift, p.norm_shift = norm_shift, p.h = h_state, p.u_in = u_in;
    p.out_f32 = out_f32, p.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16), p.out_norm = out_norm, p.aux = aux;
    {   // weights [27 * nchunk][Cout][64] bf16
        const unsigned long long dims[3] = {static_cast<unsigned long long>(BK), static_cast<unsigned long long>(Cout),
                                            static_cast<unsigned long long>(27 * p.nchunk)};
        const unsigned long long strides[2] = {2ull * BK, 2ull * BK * Cout};
        const unsigned box[3] = {BK, static_cast<unsigned>(Cout), 1};
        if (int e = encode_tensor_map(fn, &maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wpack, dims, strides, box,
                                      CU_TENSOR_MAP_SWIZZLE_128B))
            return e;
    }
    const int sms = current_sm_count(fn);
    if (sms <= 0) return 1;
    const int total = (D / TZ) * (H / TY) * (W / TX) * B;
    int ctas = total < sms ? total : sms;
    if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (Cout == 256) {
        if (int e = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(conv3d_tc_kernel<256>), Cfg<256>::SMEM)) return e;
        conv3d_tc_kernel<256><<<ctas, kThreads, Cfg<256>::SMEM, st>>>(maps, p);
    } else if (Cout == 32) {
        if (int e = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(conv3d_tc_kernel<32>), Cfg<32>::SMEM)) return e;
        conv3d_tc_kernel<32><<<ctas, kThreads, Cfg<32>::SMEM, st>>>(maps, p);
    } else {
        if (int e = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(conv3d_tc_kernel<128>), Cfg<128>::SMEM)) return e;
        conv3d_tc_kernel<128><<<ctas, kThreads, Cfg<128>::SMEM, st>>>(maps, p);
    }
    return check_launch(fn);
}

extern "C" int forge_conv3d_c8_to_1_relu(const float* x, const float* w, float bias, float* y, int B, int D, int H, int W,
                                         void* stream) {
    FORGE_RANGE("forge_conv3d_c8_to_1_relu");
    using namespace forge;
    const char* fn = "forge_conv3d_c8_to_1_relu";
    if (!x || !w || !y) return fail(fn, "null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(fn, "non-positive size");
    if (!aligned16(x)) return fail(fn, "x must be 16-byte aligned");
    const int sms = current_sm_count(fn);
    if (sms <= 0) return 1;
    const long long total = static_cast<long long>(B) * D * H * W;
    const long long want = (total + 255) / 256;
    const int grid = static_cast<int>(want < sms * 16ll ? want : sms * 16ll);
    c3d::conv_c8_to_1_relu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(x), w, bias, y, B,
                                                                                      D, H, W);
    return check_launch(fn);
}
