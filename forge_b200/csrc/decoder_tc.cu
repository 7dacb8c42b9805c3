// Tensor-core decoder: relu(conv_rgb(x)) of reference models/volume_render.py:29-37,73 with bf16
// operands / fp32 accumulation on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM)
// -- the "bf16 decoder" of BASELINE.json configs[2].  Inference only (BatchNorm in eval mode).
//
// Every layer is an implicit GEMM with NO im2col.  Activations sit in shared memory as planes of
// 16-byte records (8 bf16 channels of one pixel), records of consecutive A-rows 16 bytes apart --
// exactly the un-swizzled K-major "core matrix" layout of a UMMA operand (8 rows x 16 B contiguous,
// SBO = 128 B to the next 8 rows, LBO = distance to the next 8 values of K).  Row m of the A
// operand is the record with flattened index m, so a filter tap is the SAME descriptor with its
// start address advanced: one tcgen05.mma (M = 128 rows, K = 16) per tap and 128 flattened rows.
// Rows that fall into halo / wrap-around positions compute garbage that is never read back (an
// MMA row only depends on its own A row, so garbage stays in its row).
//
// The operand reads of the MMAs (4 KB of A + B per instruction through the shared-memory data
// pipe) are what bounds this kernel (ncu: l1tex__data_pipe_tc_wavefronts), so the plan maximises N:
//
//   layer 1  ConvTranspose2d(16->16, k6, s2, p2) = four 3x3 convolutions (one per output parity
//            class) over the SAME 9 input shifts: N = 64 = (class, co); rows = input pixels
//            (pitch 22): 3 M-tiles x 9 MMAs.
//   layer 2  Conv2d(16->8, k5).  Layer-1 pixels are stored in 8 PHASE planes by column
//            (lc % 8, record index = row * 5 + lc / 8), so one A row = a group of 8 horizontally
//            adjacent output pixels: N = 64 = (delta, co), and source column j = delta + kx
//            (0..11) is plane j % 8 shifted by j / 8 records: 1 M-tile x 5 ky x 12 columns = 60 MMAs.
//            The B operand of column j holds W2[ky][kx = j - delta] in its delta-th 8-row block; all 12
//            are windows (start block 12 - j, SBO = 128 B) into ONE strip [8 x zero, W(4), ..., W(0)].
//   layer 3  Conv2d(8->3, k5), same phase-plane scheme on the layer-2 planes (one 8-channel
//            record per pixel); K = 16 covers the rows ky and ky + 1 (A: LBO = 5 records = one row
//            down; B: LBO = one strip): 1 M-tile x 3 ky-pairs x 12 columns = 36 MMAs, N = 64 =
//            (delta, co padded to 8).
//
// The BN scales are folded into the bf16 weights; epilogues (tcgen05.ld, one accumulator row per
// thread, the two warp groups split the columns) add the shift (folded BN + bias), apply
// LeakyReLU and the zero padding of the next convolution, and
// write bf16 phase planes for the next layer; the last one applies ReLU and stores fp32 NCHW as
// float4.  Weights (46 KB of prebuilt B tiles / strips) arrive by one bulk async copy (TMA engine).
// Persistent CTAs, 2 per SM (104 KB smem, 256 TMEM columns each).
#include <cuda_bf16.h>

#include "common.cuh"

namespace forge {
namespace dtc {

constexpr int kThreads = 256;
constexpr int TOX = 32, TOY = 20;                 // output tile (20 rows: layer 2 fills 120 of its 128 MMA rows)
constexpr int IN_W = TOX / 2 + 6, IN_H = TOY / 2 + 6;     // 22 x 16 input pixels (pitch 22)
constexpr int L1_W = TOX + 8;                     // 40 x 28 layer-1 pixels = 5 groups of 8 per row
constexpr int GP = L1_W / 8;                      // record pitch of the phase planes (groups per row)
constexpr int M1_TILES = 3;                       // 128-row tiles of flattened input pixels in layer 1
constexpr int IN_PLANE = 432;                     // records per input channel-chunk plane >= 3*128 + 2*22 + 2
constexpr int L1_PLANE = 152;                     // records per layer-1 plane >= 128 + 4*5 + 1
constexpr int L2_PLANE = 160;                     // records per layer-2 plane >= 128 + 5*5 + 1
static_assert(L1_W % 8 == 0 && GP == 5, "phase planes assume 8-pixel groups");
static_assert(IN_PLANE >= M1_TILES * 128 + 2 * IN_W + 2, "input plane too small for the flattened over-read");
static_assert(IN_H * IN_W <= IN_PLANE && (TOY + 8) * GP <= L1_PLANE && (TOY + 4) * GP <= L2_PLANE && TOY % 2 == 0, "planes too small for the tile");
static_assert(L1_PLANE >= 128 + 4 * GP + 1 && L2_PLANE >= 128 + 5 * GP + 1, "phase plane too small for the over-read");
static_assert((TOY / 2 + 4) * IN_W <= M1_TILES * 128 && (TOY + 4) * GP <= 128 && TOY * GP <= 128, "M tiles do not cover the layer");

constexpr int BLK = 128;                          // one 8 x 8 bf16 core matrix
constexpr int STRIP = 13 * BLK;                   // [8 x zero][W(kx=4) .. W(kx=0)]
constexpr int W1_TILE = 2048;                     // [k/8][n = 64][k%8]
constexpr int W1_OFF = 0, W2_OFF = W1_OFF + 9 * W1_TILE, W3_OFF = W2_OFF + 10 * STRIP + 8 * BLK,
              PRM_OFF = W3_OFF + 6 * STRIP + 8 * BLK;
constexpr int WPACK_BYTES = PRM_OFF + 256;        // 47360
// the input planes alias the layer-2 planes: the former are dead once layer 1's MMAs have completed
constexpr int SM_W = 0, SM_L1 = SM_W + WPACK_BYTES, SM_L2 = SM_L1 + 16 * L1_PLANE * 16, SM_IN = SM_L2,
              SM_BAR = SM_L2 + 8 * L2_PLANE * 16, SM_TOTAL = SM_BAR + 64;
static_assert(2 * IN_PLANE * 16 <= 8 * L2_PLANE * 16, "input planes must fit in the layer-2 region they alias");
constexpr int TMEM_COLS = 256;                    // layer 1: 3 x 64 columns, layer 2: 64 (192..255), layer 3: 64 (0..63)

// instruction descriptor (PTX ISA "Instruction descriptor", kind::f16): D = f32, A = B = bf16, both K-major,
// N = 64, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_PROBE = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// shared-memory matrix descriptor, no swizzle: start address, leading (K) and stride (M/N) byte offsets
// in 16-byte units, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
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
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
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

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; ++spins) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spins > (1 << 22)) __trap();          // a lost completion must fail loudly, not hang the GPU
    }
    __syncwarp();                                 // the tcgen05.ld that follows is .sync.aligned
}
__device__ __forceinline__ void bulk_load_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.01f * v; }

// input tile: IN_H * IN_W pixels x 2 channel halves = 704 (pixel, half) items, 3 per thread (the last one partial)
constexpr int P0_ITEMS = IN_H * IN_W * 2, P0_PER_THREAD = (P0_ITEMS + kThreads - 1) / kThreads;

__device__ __forceinline__ const float4* p0_src(const float4* xin, int e, int iy0, int ix0, int Sh, int Sw) {
    const int px = e >> 1, half = e & 1;
    const int r = px / IN_W, c = px - r * IN_W;
    const int iy = iy0 + r, ix = ix0 + c;
    if (e >= P0_ITEMS || iy < 0 || iy >= Sh || ix < 0 || ix >= Sw) return nullptr;
    return xin + (static_cast<long long>(iy) * Sw + ix) * 4 + half * 2;
}

// (acc + bias) -> LeakyReLU(0.01) -> bf16 pair; max(v, 0.01 v) is LeakyReLU for a slope below 1
__device__ __forceinline__ uint32_t act_pack(float a0, float a1, float b0, float b1) {
    const float v0 = a0 + b0, v1 = a1 + b1;
    return pack_bf16(fmaxf(v0, 0.01f * v0), fmaxf(v1, 0.01f * v1));
}

__global__ void __launch_bounds__(kThreads, 2)
decoder_tc_kernel(const float* __restrict__ x, const unsigned char* __restrict__ wpack, float* __restrict__ rgb,
                  unsigned* __restrict__ masks, int Sh, int Sw, int tiles_x, int tiles_y, int total_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sW = smem + SM_W;
    unsigned char* sIn = smem + SM_IN;
    unsigned char* sL1 = smem + SM_L1;
    unsigned char* sL2 = smem + SM_L2;
    unsigned long long* wbar = reinterpret_cast<unsigned long long*>(smem + SM_BAR);
    unsigned long long* mbar = wbar + 1;                                  // [0..2]: layer-1 M-tiles, [3]: layer 2, [4]: layer 3
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 48);
    const float4* prm4 = reinterpret_cast<const float4*>(sW + PRM_OFF);   // b1[16] b2[8] b3[4]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int OH = 2 * Sh, OW = 2 * Sw;

    // ---- one-time setup: barriers, weights (bulk async copy), TMEM, zeroed activation planes ----
    if (tid == 0) {
        mbar_init(wbar, 1);
#pragma unroll
        for (int i = 0; i < 5; ++i) mbar_init(mbar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) bulk_load_g2s(sW, wpack, WPACK_BYTES, wbar);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < (SM_BAR - SM_L1) / 16; e += kThreads)          // over-read pads must hold finite values
        reinterpret_cast<uint4*>(sL1)[e] = make_uint4(0u, 0u, 0u, 0u);
    mbar_wait(wbar, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int row = (warp & 3) * 32 + lane, grp = warp >> 2;             // accumulator row of this thread, warp group
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * 32;   // lane quarter, column half
    const uint32_t aIn = smem_u32(sIn), aL1 = smem_u32(sL1), aL2 = smem_u32(sL2), aW = smem_u32(sW);
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / (tiles_x * tiles_y), t2 = tile - n * tiles_x * tiles_y;
        const int tyi = t2 / tiles_x, txi = t2 - tyi * tiles_x;
        const int Y0 = tyi * TOY, X0 = txi * TOX;

        // ---- P0: input tile fp32 NHWC -> two bf16 channel-chunk planes (zeros outside the image); all of a
        //      thread's loads are issued before the first use ----
        {
            const int iy0 = Y0 / 2 - 3, ix0 = X0 / 2 - 3;
            const float4* xin = reinterpret_cast<const float4*>(x) + static_cast<long long>(n) * Sh * Sw * 4;
            float4 v[P0_PER_THREAD][2];
#pragma unroll
            for (int k = 0; k < P0_PER_THREAD; ++k) {
                const float4* p = p0_src(xin, tid + k * kThreads, iy0, ix0, Sh, Sw);
                v[k][0] = v[k][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p) {
                    v[k][0] = __ldg(p);
                    v[k][1] = __ldg(p + 1);
                }
            }
#pragma unroll
            for (int k = 0; k < P0_PER_THREAD; ++k) {
                const int e = tid + k * kThreads;
                if (e < P0_ITEMS)
                    *reinterpret_cast<uint4*>(sIn + ((e & 1) * IN_PLANE + (e >> 1)) * 16) =
                        make_uint4(pack_bf16(v[k][0].x, v[k][0].y), pack_bf16(v[k][0].z, v[k][0].w),
                                   pack_bf16(v[k][1].x, v[k][1].y), pack_bf16(v[k][1].z, v[k][1].w));
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();

        // ---- P1: transposed conv: 3 M-tiles x 9 input shifts, N = (parity class, co); one commit per M-tile so
        //      that the epilogue of tile j overlaps the MMAs of the tiles behind it ----
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < M1_TILES; ++j) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const int a = t / 3, b = t - a * 3;       // source pixel offset (rows, cols) inside the input tile
                        umma_bf16(tmem + j * 64, make_desc(aIn + (j * 128 + a * IN_W + b) * 16, IN_PLANE * 16, 128),
                                  make_desc(aW + W1_OFF + t * W1_TILE, 1024, 128), IDESC, t > 0);
                    }
                    umma_commit(mbar + j);
                }
            }
            __syncwarp();
        }
        // the next tile's input is pulled into L2 while this tile computes
        {
            const int nt = tile + gridDim.x;
            if (nt < total_tiles) {
                const int nn = nt / (tiles_x * tiles_y), nt2 = nt - nn * tiles_x * tiles_y;
                const int nty = nt2 / tiles_x, ntx = nt2 - nty * tiles_x;
                const float4* xin = reinterpret_cast<const float4*>(x) + static_cast<long long>(nn) * Sh * Sw * 4;
#pragma unroll
                for (int k = 0; k < P0_PER_THREAD; ++k) {
                    const float4* p = p0_src(xin, tid + k * kThreads, nty * TOY / 2 - 3, ntx * TOX / 2 - 3, Sh, Sw);
                    if (p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
        }
        {
            const float4 b1a = prm4[0], b1b = prm4[1], b1c = prm4[2], b1d = prm4[3];
#pragma unroll 1
            for (int j = 0; j < M1_TILES; ++j) {
                mbar_wait(mbar + j, phase);
                tc_fence_after();
                float acc[32];                                    // classes (py = grp, px = 0) and (py = grp, px = 1)
                tmem_ld32(t_lane + j * 64, acc);
                const int m = j * 128 + row, yr = m / IN_W, xr = m - yr * IN_W;
                if (yr < TOY / 2 + 4 && xr < TOX / 2 + 4) {
                    const int lr = 2 * yr + grp, oy = Y0 - 4 + lr;
#pragma unroll
                    for (int px = 0; px < 2; ++px) {
                        const int lc = 2 * xr + px, ox = X0 - 4 + lc;
                        const float* a = acc + px * 16;
                        uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
                        if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) {       // else: zero padding of layer 2
                            q0 = make_uint4(act_pack(a[0], a[1], b1a.x, b1a.y), act_pack(a[2], a[3], b1a.z, b1a.w),
                                            act_pack(a[4], a[5], b1b.x, b1b.y), act_pack(a[6], a[7], b1b.z, b1b.w));
                            q1 = make_uint4(act_pack(a[8], a[9], b1c.x, b1c.y), act_pack(a[10], a[11], b1c.z, b1c.w),
                                            act_pack(a[12], a[13], b1d.x, b1d.y), act_pack(a[14], a[15], b1d.z, b1d.w));
                        }
                        unsigned char* dst = sL1 + (((lc & 7) * 2) * L1_PLANE + lr * GP + (lc >> 3)) * 16;
                        *reinterpret_cast<uint4*>(dst) = q0;
                        *reinterpret_cast<uint4*>(dst + L1_PLANE * 16) = q1;
                        if (masks && lr >= 4 && lr < 4 + TOY && lc >= 4 && lc < 4 + TOX && oy < OH && ox < OW) {
                            // sign bits of the 16 pre-activations for the backward pass (decoder_bwd.cu), low half-word
                            const float bb[16] = {b1a.x, b1a.y, b1a.z, b1a.w, b1b.x, b1b.y, b1b.z, b1b.w,
                                                  b1c.x, b1c.y, b1c.z, b1c.w, b1d.x, b1d.y, b1d.z, b1d.w};
                            unsigned bits = 0;
#pragma unroll
                            for (int c = 0; c < 16; ++c) bits |= (a[c] + bb[c] > 0.f ? 1u : 0u) << c;
                            reinterpret_cast<unsigned short*>(masks + (static_cast<long long>(n) * OH + oy) * OW + ox)[0] =
                                static_cast<unsigned short>(bits);
                        }
                    }
                }
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();

        // ---- P2: conv 5x5 16 -> 8: rows = groups of 8 pixels, 5 ky x 12 source columns ----
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        umma_bf16(tmem + 192,
                                  make_desc(aL1 + (((j & 7) * 2) * L1_PLANE + ky * GP + (j >> 3)) * 16, L1_PLANE * 16, 128),
                                  make_desc(aW + W2_OFF + ky * 2 * STRIP + (12 - j) * BLK, STRIP, 128), IDESC, (ky | j) > 0);
                umma_commit(mbar + 3);
            }
            __syncwarp();
        }
        {
            const float4 b2a = prm4[4], b2b = prm4[5];
            mbar_wait(mbar + 3, phase);
            tc_fence_after();
            float acc[32];                                        // pixels delta = 4 grp .. 4 grp + 3 of this row's group
            tmem_ld32(t_lane + 192, acc);
            const int yr = row / GP, xg = row - yr * GP, oy = Y0 - 2 + yr;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int delta = 4 * grp + d, ox = X0 - 2 + 8 * xg + delta;
                const float* a = acc + d * 8;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (oy >= 0 && oy < OH && ox >= 0 && ox < OW)                  // else: zero padding of layer 3
                    q = make_uint4(act_pack(a[0], a[1], b2a.x, b2a.y), act_pack(a[2], a[3], b2a.z, b2a.w),
                                   act_pack(a[4], a[5], b2b.x, b2b.y), act_pack(a[6], a[7], b2b.z, b2b.w));
                *reinterpret_cast<uint4*>(sL2 + (delta * L2_PLANE + row) * 16) = q;
                const int xl = 8 * xg + delta;
                if (masks && yr >= 2 && yr < 2 + TOY && xl >= 2 && xl < 2 + TOX && oy < OH && ox < OW) {
                    const float bb[8] = {b2a.x, b2a.y, b2a.z, b2a.w, b2b.x, b2b.y, b2b.z, b2b.w};
                    unsigned bits = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) bits |= (a[c] + bb[c] > 0.f ? 1u : 0u) << c;
                    reinterpret_cast<unsigned char*>(masks + (static_cast<long long>(n) * OH + oy) * OW + ox)[2] =
                        static_cast<unsigned char>(bits);
                }
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();

        // ---- P3: conv 5x5 8 -> 3: K = rows (ky, ky + 1), 3 ky pairs x 12 source columns ----
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int kp = 0; kp < 3; ++kp)
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        umma_bf16(tmem, make_desc(aL2 + ((j & 7) * L2_PLANE + 2 * kp * GP + (j >> 3)) * 16, GP * 16, 128),
                                  make_desc(aW + W3_OFF + 2 * kp * STRIP + (12 - j) * BLK, STRIP, 128), IDESC, (kp | j) > 0);
                umma_commit(mbar + 4);
            }
            __syncwarp();
        }
        {
            const float4 b3 = prm4[6];
            mbar_wait(mbar + 4, phase);
            tc_fence_after();
            float acc[32];
            tmem_ld32(t_lane, acc);
            const int yr = row / GP, xg = row - yr * GP;
            const int oy = Y0 + yr, ox = X0 + 8 * xg + 4 * grp;
            if (yr < TOY && xg < TOX / 8 && oy < OH && ox < OW) {
                const float bias[3] = {b3.x, b3.y, b3.z};
                if (masks) {
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        if (ox + d < OW) {
                            const unsigned bits = (acc[8 * d] + b3.x > 0.f ? 1u : 0u) | (acc[8 * d + 1] + b3.y > 0.f ? 2u : 0u) |
                                                  (acc[8 * d + 2] + b3.z > 0.f ? 4u : 0u);
                            reinterpret_cast<unsigned char*>(masks + (static_cast<long long>(n) * OH + oy) * OW + ox + d)[3] =
                                static_cast<unsigned char>(bits);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float b = bias[c];
                    float* dst = rgb + ((static_cast<long long>(n) * 3 + c) * OH + oy) * OW + ox;
                    const float4 v = make_float4(fmaxf(acc[c] + b, 0.f), fmaxf(acc[8 + c] + b, 0.f), fmaxf(acc[16 + c] + b, 0.f),
                                                 fmaxf(acc[24 + c] + b, 0.f));
                    if (ox + 3 < OW && (OW & 3) == 0) {
                        __stcs(reinterpret_cast<float4*>(dst), v);
                    } else {
                        dst[0] = v.x;
                        if (ox + 1 < OW) dst[1] = v.y;
                        if (ox + 2 < OW) dst[2] = v.z;
                        if (ox + 3 < OW) dst[3] = v.w;
                    }
                }
            }
        }
        phase ^= 1;                                               // every barrier completed exactly once per tile
        tc_fence_before();
        __syncthreads();
    }

    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// Test hook kernel: ONE tcgen05.mma (M=128, N=16, K=16, bf16) over a caller-provided shared-memory image with
// caller-provided descriptor fields; D [128][16] fp32 comes back through tcgen05.ld.
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const unsigned char* __restrict__ image, int image_bytes, unsigned a_off, unsigned a_lbo, unsigned a_sbo,
                  unsigned b_off, unsigned b_lbo, unsigned b_sbo, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < image_bytes / 16; e += 128) reinterpret_cast<uint4*>(smem)[e] = reinterpret_cast<const uint4*>(image)[e];
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        umma_bf16(tmem, make_desc(smem_u32(smem) + a_off, a_lbo, a_sbo), make_desc(smem_u32(smem) + b_off, b_lbo, b_sbo), IDESC_PROBE, 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float acc[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
#pragma unroll
    for (int c = 0; c < 16; ++c) out[(warp * 32 + lane) * 16 + c] = acc[c];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
    }
}

}  // namespace dtc
}  // namespace forge

extern "C" int forge_decoder_tc_wpack_bytes(void) { return forge::dtc::WPACK_BYTES; }

extern "C" int forge_decoder_tc_fwd(const float* x_nhwc, const void* wpack, float* rgb_nchw, unsigned* sign_masks, int N,
                                    int S_h, int S_w, int max_ctas, void* stream) {
    FORGE_RANGE("forge_decoder_tc_fwd");
    using namespace forge;
    using namespace forge::dtc;
    const char* fn = "forge_decoder_tc_fwd";
    if (!x_nhwc || !wpack || !rgb_nchw) return fail(fn, "null pointer");
    if (N <= 0 || S_h <= 0 || S_w <= 0) return fail(fn, "non-positive size");
    if (!aligned16(x_nhwc) || !aligned16(wpack)) return fail(fn, "x_nhwc / wpack must be 16-byte aligned");
    const int tiles_x = (2 * S_w + TOX - 1) / TOX, tiles_y = (2 * S_h + TOY - 1) / TOY;
    const long long total = static_cast<long long>(N) * tiles_x * tiles_y;
    if (total > 0x7fffffffLL) return fail(fn, "too many tiles for one launch");
    if (int rc = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(decoder_tc_kernel), SM_TOTAL)) return rc;
    const int sm_count = current_sm_count(fn);
    if (sm_count <= 0) return 1;
    int ctas = 2 * sm_count;                       // persistent: two resident CTAs per SM
    if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
    if (total < ctas) ctas = static_cast<int>(total);
    decoder_tc_kernel<<<ctas, kThreads, SM_TOTAL, static_cast<cudaStream_t>(stream)>>>(
        x_nhwc, static_cast<const unsigned char*>(wpack), rgb_nchw, sign_masks, S_h, S_w, tiles_x, tiles_y,
        static_cast<int>(total));
    return check_launch(fn);
}

extern "C" int forge_umma_probe(const void* image, int image_bytes, unsigned a_off, unsigned a_lbo, unsigned a_sbo,
                                unsigned b_off, unsigned b_lbo, unsigned b_sbo, float* out, void* stream) {
    FORGE_RANGE("forge_umma_probe");
    using namespace forge;
    using namespace forge::dtc;
    const char* fn = "forge_umma_probe";
    if (!image || !out) return fail(fn, "null pointer");
    if (image_bytes <= 0 || image_bytes % 16 || image_bytes > 200 * 1024) return fail(fn, "image_bytes must be a multiple of 16, <= 200 KiB");
    if ((a_off | a_lbo | a_sbo | b_off | b_lbo | b_sbo) & 15u) return fail(fn, "offsets must be multiples of 16 bytes");
    if (int rc = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(umma_probe_kernel), image_bytes)) return rc;
    umma_probe_kernel<<<1, 128, image_bytes, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned char*>(image), image_bytes, a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, out);
    return check_launch(fn);
}
