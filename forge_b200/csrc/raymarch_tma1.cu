// K1-T1: TMA-staged voxel bricks with ONE lane per ray (raymarch_tma.cu maps two lanes to a ray).
//
// Why: after the brick staging K1 is bound by instruction issue, not by the data pipe, and ~100 of the ~230 SASS instructions
// of a warp iteration are per-SAMPLE overhead (position, floor, weights, density, in-brick test) that does not depend on the
// channel split.  One lane per ray amortises that overhead over 32 rays per warp instruction instead of 16: ~8.6 instead of
// ~14.4 issued instructions per ray-sample.
//
// Keeping the shared-memory reads conflict-free with 8 DIFFERENT rays per LDS.128 phase:
//   * a phase (8 consecutive lanes) is a 4 x 2 pixel block; lane l reads the 16-byte chunk (i ^ (l & 3)) of its voxel, so the
//     four lanes of a pixel row use four different chunk slots;
//   * the two pixel rows of a phase (lanes l and l + 4) share a slot, and their base voxels are almost always identical (a
//     broadcast) or face neighbours.  The brick is stored with ODD pitches (in-plane shapes 9..17 x 9..17 voxels, planes
//     fetched in pairs so that every TMA destination stays 128-byte aligned): the bank half of a voxel is then the parity of
//     x + y + z, and face neighbours always sit in opposite halves.  Simulated on the cfg-2 cameras: 1.05 wavefronts per phase
//     (1.49 with even pitches, 1.85 with 8 x 1 pixel phases).
//
// Everything else -- slab planning by a producer warp, full / empty mbarrier ring, exact fallback to direct gathers for
// samples outside the resident brick, density software-pipelined one sample ahead -- is the scheme of raymarch_tma.cu.
// CTA = 16 x 16 pixel tile: 8 consumer warps (8 x 4 pixels each) + the producer; two CTAs per SM.
#include <cstdlib>

#include "raymarch_tma_common.cuh"

namespace forge {

constexpr int kNumB1 = 5;
__constant__ int c_box1[kNumB1] = {9, 11, 13, 15, 17};
constexpr int kBox1[kNumB1] = {9, 11, 13, 15, 17};
struct TmaMaps1 {
    CUtensorMap m[kNumB1 * kNumB1];      // index = iy * kNumB1 + ix; box = (16 ch, bx, by, 2 planes)
};

constexpr int kT1 = 16;                  // pixel tile side
constexpr int kCons1 = 8;                // consumer warps
constexpr int kThreads1 = 32 * (kCons1 + 1);

template <int kStages, int kStageVox>
__global__ void __launch_bounds__(kThreads1, 2)
raymarch_fwd_tma1_kernel(const __grid_constant__ TmaMaps1 maps, const float* __restrict__ feat_pad,
                         const float4* __restrict__ dens_quad, const int* __restrict__ view2vol,
                         const float* __restrict__ cam12, const float* __restrict__ zs_g, float* __restrict__ out_feat,
                         float* __restrict__ out_sil, float* __restrict__ out_depth, int D, int H, int W, int Sh, int Sw, int P,
                         int tiles_x, int interleave) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw + kStages * kStageVox * 64);
    const uint32_t stage0 = smem_u32(smem_raw);

    // heavy-first schedule: views interleaved, tiles ranked centre-out
    const int order = blockIdx.y * gridDim.x + blockIdx.x, n_views = gridDim.y, tiles_y = gridDim.x / tiles_x;
    const int rank = interleave ? order / n_views : static_cast<int>(blockIdx.x);
    const int n = interleave ? order - rank * n_views : static_cast<int>(blockIdx.y);
    const int ty = centre_out(rank / tiles_x, tiles_y), tx = centre_out(rank % tiles_x, tiles_x);

    for (int k = threadIdx.x; k < P; k += kThreads1) sm.zs[k] = zs_g[k];
    if (threadIdx.x < 12) sm.cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kCons1);
        }
        sm.kt0 = P;
        sm.kt1 = 0;
        mbar_init_fence();
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wp = W + 2, Hp = H + 2, Wq = W + 1, Hq = H + 1;
    const long long v = view2vol[n];
    const float* fvol = feat_pad + v * (D + 2) * Hp * Wp * 16;

    // ---- per-ray setup: warp = 8 x 4 pixels, phase (8 lanes) = 4 x 2 pixels ----
    const int ph = lane >> 3, l8 = lane & 7, rq = l8 & 3;
    const int j = tx * kT1 + (warp & 1) * 8 + (ph & 1) * 4 + (l8 & 3);
    const int i = ty * kT1 + ((warp >> 1) & 3) * 4 + (ph >> 1) * 2 + (l8 >> 2);
    const bool valid = (warp < kCons1) && (i < Sh) && (j < Sw);
    Ray r;
    r.k0 = r.k1 = 0;
    if (warp < kCons1) r = make_ray(sm.cam, i, j, sm.zs, P, D, H, W);
    if (!valid) r.k1 = 0;
    int kw0 = r.k1 > r.k0 ? r.k0 : P, kw1 = r.k1 > r.k0 ? r.k1 : 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        kw0 = min(kw0, __shfl_xor_sync(0xffffffffu, kw0, s));
        kw1 = max(kw1, __shfl_xor_sync(0xffffffffu, kw1, s));
    }
    if (lane == 0 && warp < kCons1 && kw1 > kw0) {
        atomicMin(&sm.kt0, kw0);
        atomicMax(&sm.kt1, kw1);
    }
    __syncthreads();
    const int kt0 = sm.kt0, kt1 = sm.kt1;
    if (kt1 <= kt0) {          // the whole tile misses the volume
        if (valid) {
            const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
            float4* o = reinterpret_cast<float4*>(out_feat + pix * 16);
            o[0] = o[1] = o[2] = o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
            out_sil[pix] = 0.f;
            if (out_depth) out_depth[pix] = 0.f;
        }
        return;
    }

    // ---- slab planning and brick copies (whole producer warp; all lanes hold the same result) ----
    const float u0 = static_cast<float>(tx * kT1) + 0.5f, u1 = static_cast<float>(min(tx * kT1 + kT1, Sw) - 1) + 0.5f;
    const float v0 = static_cast<float>(ty * kT1) + 0.5f, v1 = static_cast<float>(min(ty * kT1 + kT1, Sh) - 1) + 0.5f;
    const int zvol = static_cast<int>(v) * (D + 2);
    auto plan = [&](int s, int ka) {
        const int cand = min(ka + 1 + min(lane, kSlabMax - 1), kt1);
        Box mine;
        mine.lo[0] = mine.lo[1] = mine.lo[2] = 0;
        mine.ex[0] = mine.ex[1] = mine.ex[2] = 0;
        if (ka < kt1) mine = slab_box(sm.cam, sm.zs, ka, cand, u0, u1, v0, v1, D, H, W);
        int ix = 0, iy = 0;                                  // smallest odd shape covering the footprint in x, y
#pragma unroll
        for (int e = 0; e < kNumB1 - 1; ++e) {
            ix += (mine.ex[0] > c_box1[e]);
            iy += (mine.ex[1] > c_box1[e]);
        }
        const int bx = c_box1[ix], by = c_box1[iy], pairs = (mine.ex[2] + 1) >> 1;
        const bool empty = mine.ex[0] == 0;
        // a candidate fits when its footprint is covered by the largest shape AND the brick fits the stage
        const bool covered = mine.ex[0] <= c_box1[kNumB1 - 1] && mine.ex[1] <= c_box1[kNumB1 - 1];
        const unsigned fits = __ballot_sync(0xffffffffu, empty || (covered && bx * by * 2 * pairs <= kStageVox)) & ((1u << kSlabMax) - 1u);
        const int pick = max(__ffs(~fits) - 1, 1) - 1;      // longest run of fitting candidates, at least one sample
        if (lane == pick) {
            SlabHeader h;
            h.ka = ka, h.kb = (ka < kt1) ? cand : ka;
            h.lx = mine.lo[0], h.ly = mine.lo[1], h.lz = mine.lo[2];
            h.ex = empty ? 0 : bx, h.ey = empty ? 0 : by;
            h.ez = empty ? 0 : 2 * min(pairs, kStageVox / (2 * bx * by));   // too deep for a stage: clip (the rest gathers)
            h.shape = iy * kNumB1 + ix;
            sm.hdr[s & 7] = h;
        }
        __syncwarp();
    };
    auto issue = [&](int s) {
        const int st = s % kStages;
        const SlabHeader h = sm.hdr[s & 7];
        const uint32_t pair_bytes = static_cast<uint32_t>(h.ex * h.ey) * 128u;
        if (lane == 0) {
            if (h.ez > 0) mbar_arrive_expect_tx(&sm.full[st], pair_bytes * static_cast<uint32_t>(h.ez >> 1));
            else mbar_arrive(&sm.full[st]);
        }
        __syncwarp();
        const uint32_t dst0 = stage0 + static_cast<uint32_t>(st) * (kStageVox * 64);
        if (lane < (h.ez >> 1))
            tma_load_4d(dst0 + static_cast<uint32_t>(lane) * pair_bytes, &maps.m[h.shape], 0, h.lx, h.ly, zvol + h.lz + 2 * lane,
                        &sm.full[st]);
    };
    if (warp == kCons1) {
        // ================= producer warp =================
        if (lane < kNumB1 * kNumB1) prefetch_tensormap(&maps.m[lane]);
        plan(0, kt0);
        for (int s = 0; sm.hdr[s & 7].ka < kt1; ++s) {
            mbar_wait(&sm.empty[s % kStages], ((s / kStages) & 1) ^ 1);
            issue(s);
            plan(s + 1, sm.hdr[s & 7].kb);
        }
        return;
    }

    // ================= consumer warps =================
    const float4* qv = dens_quad + v * (D + 2) * Hq * Wq;
    const int row_y = Wp * 16, row_z = Hp * Wp * 16;          // float strides of the padded feature volume
    uint32_t choff[4];                                        // byte offset of the chunk this lane reads i-th
#pragma unroll
    for (int e = 0; e < 4; ++e) choff[e] = static_cast<uint32_t>((e ^ rq) << 4);

    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    float T = 1.f, depth = 0.f;

    // software pipeline, one sample ahead: footprint + both density quads of sample k + 1 before the feature work of sample k
    auto fetch = [&](int kk, Foot& f, float4& q0, float4& q1) -> bool {
        f = sample_foot(r, sm.zs[kk], D, H, W);
        const bool act = f.in && (kk >= r.k0) && (kk < r.k1);
        if (act) {
            const float4* q = qv + (static_cast<long long>(f.z0 + 1) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1);
            q0 = __ldg(q);
            q1 = __ldg(q + Hq * Wq);
        }
        return act;
    };
    Foot fn;
    float4 q0n = make_float4(0.f, 0.f, 0.f, 0.f), q1n = q0n;
    bool actn = false;
    if (kw1 > kw0) actn = fetch(max(kt0, kw0), fn, q0n, q1n);

    for (int s = 0;; ++s) {
        const int st = s % kStages;
        mbar_wait(&sm.full[st], (s / kStages) & 1);
        const SlabHeader h = sm.hdr[s & 7];
        const uint32_t brick = stage0 + static_cast<uint32_t>(st) * (kStageVox * 64);
        const uint32_t sy = static_cast<uint32_t>(h.ex) << 6, sz = static_cast<uint32_t>(h.ex * h.ey) << 6;
        const int kend = min(h.kb, kw1);
        for (int k = max(h.ka, kw0); k < kend; ++k) {
            const float z = sm.zs[k];
            const Foot f = fn;
            const bool act = actn;
            const float4 q0 = q0n, q1 = q1n;
            if (k + 1 < kw1) actn = fetch(k + 1, fn, q0n, q1n);
            // ATen weight order: (wx * wy) * wz
            const float w00 = __fmul_rn(f.wx0, f.wy0), w10 = __fmul_rn(f.wx1, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1),
                        w11 = __fmul_rn(f.wx1, f.wy1);
            float sigma = 0.f;
            if (act) {          // same per-plane partial sums as the two-lane kernels, added in the same order
                float p0 = __fmul_rn(w00, f.wz0) * q0.x;
                p0 = fmaf(__fmul_rn(w10, f.wz0), q0.y, p0);
                p0 = fmaf(__fmul_rn(w01, f.wz0), q0.z, p0);
                p0 = fmaf(__fmul_rn(w11, f.wz0), q0.w, p0);
                float p1 = __fmul_rn(w00, f.wz1) * q1.x;
                p1 = fmaf(__fmul_rn(w10, f.wz1), q1.y, p1);
                p1 = fmaf(__fmul_rn(w01, f.wz1), q1.z, p1);
                p1 = fmaf(__fmul_rn(w11, f.wz1), q1.w, p1);
                sigma = p0 + p1;
            }
            const float wk = sigma * T;
            if (wk != 0.f) {    // sigma != 0 implies act
                const float wxy[4] = {w00, w10, w01, w11};
                const int xb = f.x0 + 1 - h.lx, yb = f.y0 + 1 - h.ly, zb = f.z0 + 1 - h.lz;
                const bool inbox = (static_cast<unsigned>(xb) + 1u < static_cast<unsigned>(h.ex)) &&
                                   (static_cast<unsigned>(yb) + 1u < static_cast<unsigned>(h.ey)) &&
                                   (static_cast<unsigned>(zb) + 1u < static_cast<unsigned>(h.ez));
                if (inbox) {
                    const uint32_t a0 = brick + static_cast<uint32_t>(((zb * h.ey + yb) * h.ex + xb) << 6);
#pragma unroll
                    for (int cn = 0; cn < 8; ++cn) {        // corner bit0 = dx, bit1 = dy, bit2 = dz
                        const float cw = wk * __fmul_rn(wxy[cn & 3], (cn & 4) ? f.wz1 : f.wz0);
                        const uint32_t a = a0 + ((cn & 1) ? 64u : 0u) + ((cn & 2) ? sy : 0u) + ((cn & 4) ? sz : 0u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw, lds128(a + choff[e]));
                    }
                } else {
                    const float* p = fvol + ((f.z0 + 1) * Hp + (f.y0 + 1)) * row_y + (f.x0 + 1) * 16;
#pragma unroll
                    for (int cn = 0; cn < 8; ++cn) {
                        const float cw = wk * __fmul_rn(wxy[cn & 3], (cn & 4) ? f.wz1 : f.wz0);
                        const float* pc = p + ((cn & 1) ? 16 : 0) + ((cn & 2) ? row_y : 0) + ((cn & 4) ? row_z : 0);
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw, ldg128(pc + (choff[e] >> 2)));
                    }
                }
                depth = fmaf(wk, z, depth);
            }
            T = T * (1.f - sigma);
        }
        if (h.kb >= kt1) break;                               // that was the last slab
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);            // release the stage to the producer
    }

    // ---- store (acc[4 e + t] is channel 4 (e ^ rq) + t) ----
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        float* o = out_feat + pix * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            *reinterpret_cast<float4*>(o + (choff[e] >> 2)) = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
        out_sil[pix] = 1.f - T;
        if (out_depth) out_depth[pix] = depth;
    }
}

int raymarch_fwd_tma1_launch(const char* fn, const float* feat_pad, const float4* dens_quad, const int* view2vol,
                             const float* cam12, const float* zs, float* out_feat, float* out_sil, float* out_depth, int N,
                             int V, int D, int H, int W, int S_h, int S_w, int P, cudaStream_t st) {
    constexpr int kStages = 2, kStageVox = 832;
    constexpr int bytes = tma_smem_bytes(kStages, kStageVox);
    static_assert(2 * bytes <= 227 * 1024, "two CTAs per SM must fit");
    static_assert(kStageVox >= 17 * 17 * 2, "a stage must hold one plane pair of the largest shape");
    alignas(64) TmaMaps1 maps;
    const unsigned long long dims[4] = {16ull, static_cast<unsigned long long>(W + 2), static_cast<unsigned long long>(H + 2),
                                        static_cast<unsigned long long>(D + 2) * V};
    const unsigned long long strides[3] = {64ull, 64ull * (W + 2), 64ull * (W + 2) * (H + 2)};
    for (int iy = 0; iy < kNumB1; ++iy)
        for (int ix = 0; ix < kNumB1; ++ix) {
            const unsigned box[4] = {16u, static_cast<unsigned>(kBox1[ix]), static_cast<unsigned>(kBox1[iy]), 2u};
            if (int e = encode_tensor_map(fn, &maps.m[iy * kNumB1 + ix], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, feat_pad, dims, strides,
                                          box, CU_TENSOR_MAP_SWIZZLE_NONE))
                return e;
        }
    if (int e = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(raymarch_fwd_tma1_kernel<kStages, kStageVox>), bytes)) return e;
    const int tiles_x = (S_w + kT1 - 1) / kT1, tiles_y = (S_h + kT1 - 1) / kT1;
    dim3 grid(tiles_x * tiles_y, N);
    raymarch_fwd_tma1_kernel<kStages, kStageVox><<<grid, kThreads1, bytes, st>>>(
        maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, D, H, W, S_h, S_w, P, tiles_x,
        interleave_views(V, D, H, W));
    return check_launch(fn);
}

}  // namespace forge
