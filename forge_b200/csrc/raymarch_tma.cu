// K1-T: the raymarch forward with TMA-staged voxel bricks in shared memory (the formulation BASELINE.json's north_star
// names).  Same inputs, outputs and sampling arithmetic as raymarch_fwd_kernel (raymarch.cu); what changes is the operand
// path of the 16-channel features:
//
//   * a CTA (8 warps) marches a 16x8 pixel tile of one view in k-SLABS (runs of consecutive sample depths).  The samples of
//     a slab form a frustum segment whose voxel footprint is bounded by the axis-aligned box of its 8 vertices (positions
//     are multilinear in pixel and depth).  The brick is fetched from the packed volume as z-planes of a static in-plane
//     shape, one cp.async.bulk.tensor (UTMALDG) per plane, into a 2-stage shared-memory ring; slab lengths adapt so that
//     every brick fits its stage (<= 880 voxels = 55 KB).
//   * a ninth warp is the producer: it waits for a stage to be released (empty mbarrier, one arrival per consumer warp),
//     issues the copies of the next slab (planned while the previous copies were in flight) and plans the one after it.
//     (Variant without a producer warp -- the last consumer warp to finish a slab refills the stage, 128 registers per
//     thread -- measured slower, 0.370 vs 0.357 ms at cfg-2: the refill lands on the critical path of the slowest warp.)
//   * corners are read with LDS.128.  Mapping: 2 lanes per ray, lane c owns the x-corner x0 + c and all 16 channels; the
//     four 16-byte chunks of a voxel are read in an order rotated by the ray's index in its quarter-warp
//     (chunk = i ^ (ray & 3)), so the 8 lanes of a shared-memory phase always hit 8 distinct 4-bank groups: x0 and x0+1
//     are adjacent 64-byte records (opposite bank halves) and the four rays use four different chunk slots.  Every
//     LDS.128 phase is one conflict-free wavefront: 4 wavefronts per ray-sample, the minimum for 512 bytes, where the
//     direct L1 gathers of raymarch_fwd_kernel need ~6.3 (one per distinct 128-byte line of a request).
//
// A sample whose footprint is not completely inside the resident brick (rounding at the box faces, clipped boxes) takes the
// direct global path, so results never depend on the box arithmetic.  Density stays on the global path (dens_quad: 32
// bytes per sample against 512 for the features), software-pipelined one sample ahead.
#include <cstdlib>

#include "raymarch_tma_common.cuh"

namespace forge {

// Static in-plane box shapes of the tensor maps (voxels in x, y): a slab's brick is fetched as ez z-planes of the smallest
// shape that covers its footprint, one cp.async.bulk.tensor (UTMALDG) per plane.  (Row-wise cp.async.bulk copies, ~53 per
// slab, saturated the SM's TMA unit at ~50 cycles per request: ncu showed the producer warp 88 % busy issuing them.)
constexpr int kNumBX = 4, kNumBY = 4;
__constant__ int c_box_x[kNumBX] = {8, 10, 12, 16};
__constant__ int c_box_y[kNumBY] = {6, 8, 12, 16};
constexpr int kBoxX[kNumBX] = {8, 10, 12, 16};
constexpr int kBoxY[kNumBY] = {6, 8, 12, 16};
struct TmaMaps {
    CUtensorMap m[kNumBX * kNumBY];      // index = iy * kNumBX + ix
};

// kStages x kStageVox voxels (64 B each) of brick ring per CTA; two CTAs per SM
template <int kStages, int kStageVox, int kWarpsY, bool kPrefetch = true, int kMode = 0>
__global__ void __launch_bounds__(32 * (4 * kWarpsY + 1), kWarpsY <= 2 ? 2 : 1)
raymarch_fwd_tma_kernel(const __grid_constant__ TmaMaps maps, const float* __restrict__ feat_pad,
                        const float4* __restrict__ dens_quad,
                        const int* __restrict__ view2vol, const float* __restrict__ cam12, const float* __restrict__ zs_g,
                        float* __restrict__ out_feat, float* __restrict__ out_sil, float* __restrict__ out_depth, int D,
                        int H, int W, int Sh, int Sw, int P, int tiles_x, int interleave) {
    constexpr int kConsWarps = 4 * kWarpsY, kTmaThreads = 32 * (kConsWarps + 1), kTH = 4 * kWarpsY;
    constexpr bool kTuned = kMode >= 1;          // 0: first round-2 sample loop, 1 / 3 / 4: tuned instruction stream with 4 / 8 / 16 LDS.128 in flight, 5: 16 in flight and one sample per loop trip (default)
    // (A sample-pair mapping -- lane c takes sample k + c and reads all 8 corners, so the footprint / density / weight work is done
    // once per ray-sample instead of once per lane -- was built and measured in commit "K1-T sample-pair mapping": same instruction
    // count, 20 more live registers under the 96-register cap of 2 x 9 warps, 51 M local-memory sectors of spill traffic through L1:
    // 0.556 vs 0.333 ms.  profiles/r02_k1t_pair.txt)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw + kStages * kStageVox * 64);
    const uint32_t stage0 = smem_u32(smem_raw);

    // heavy-first schedule, as in raymarch_fwd_kernel: views interleaved, tiles ranked centre-out
    const int order = blockIdx.y * gridDim.x + blockIdx.x, n_views = gridDim.y, tiles_y = gridDim.x / tiles_x;
    const int rank = interleave ? order / n_views : static_cast<int>(blockIdx.x);
    const int n = interleave ? order - rank * n_views : static_cast<int>(blockIdx.y);
    const int ty = centre_out(rank / tiles_x, tiles_y), tx = centre_out(rank % tiles_x, tiles_x);

    for (int k = threadIdx.x; k < P; k += kTmaThreads) sm.zs[k] = zs_g[k];
    if (threadIdx.x < 12) sm.cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kConsWarps);
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

    // ---- per-ray setup and the tile's sample range ----
    const int c = lane & 1, q = lane >> 1, rq = q & 3;
    const int j = tx * kTW + (warp & 3) * 4 + (q & 3);
    const int i = ty * kTH + ((warp >> 2) % kWarpsY) * 4 + (q >> 2);
    const bool valid = (warp < kConsWarps) && (i < Sh) && (j < Sw);
    Ray r;
    r.k0 = r.k1 = 0;
    if (warp < kConsWarps) r = make_ray(sm.cam, i, j, sm.zs, P, D, H, W);
    if (!valid) r.k1 = 0;
    int kw0 = r.k1 > r.k0 ? r.k0 : P, kw1 = r.k1 > r.k0 ? r.k1 : 0;
#pragma unroll
    for (int s = 16; s >= 2; s >>= 1) {
        kw0 = min(kw0, __shfl_xor_sync(0xffffffffu, kw0, s));
        kw1 = max(kw1, __shfl_xor_sync(0xffffffffu, kw1, s));
    }
    if (lane == 0 && warp < kConsWarps && kw1 > kw0) {
        atomicMin(&sm.kt0, kw0);
        atomicMax(&sm.kt1, kw1);
    }
    __syncthreads();
    const int kt0 = sm.kt0, kt1 = sm.kt1;
    if (kt1 <= kt0) {          // the whole tile misses the volume
        if (valid) {
            const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
            float4* o = reinterpret_cast<float4*>(out_feat + pix * 16 + c * 8);
            o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c == 0) {
                out_sil[pix] = 0.f;
                if (out_depth) out_depth[pix] = 0.f;
            }
        }
        return;
    }

    // ---- slab planning and brick copies (executed by whole warps; all lanes hold the same result) ----
    const float u0 = static_cast<float>(tx * kTW) + 0.5f, u1 = static_cast<float>(min(tx * kTW + kTW, Sw) - 1) + 0.5f;
    const float v0 = static_cast<float>(ty * kTH) + 0.5f, v1 = static_cast<float>(min(ty * kTH + kTH, Sh) - 1) + 0.5f;
    const int zvol = static_cast<int>(v) * (D + 2);
    // plan slab s starting at sample ka into hdr[s & 7]: lane L sizes the box of samples [ka, ka + 1 + L); the longest run
    // whose brick (static in-plane shape x its z extent) fits a stage wins.  ka >= kt1 plans an empty end marker.
    auto plan = [&](int s, int ka) {
        const int cand = min(ka + 1 + min(lane, kSlabMax - 1), kt1);
        Box mine;
        mine.lo[0] = mine.lo[1] = mine.lo[2] = 0;
        mine.ex[0] = mine.ex[1] = mine.ex[2] = 0;
        if (ka < kt1) mine = slab_box(sm.cam, sm.zs, ka, cand, u0, u1, v0, v1, D, H, W);
        int ix = 0, iy = 0;                                  // smallest static shape covering the footprint in x, y
#pragma unroll
        for (int e = 0; e < kNumBX - 1; ++e) ix += (mine.ex[0] > c_box_x[e]);
#pragma unroll
        for (int e = 0; e < kNumBY - 1; ++e) iy += (mine.ex[1] > c_box_y[e]);
        const int bx = c_box_x[ix], by = c_box_y[iy];
        const bool empty = mine.ex[0] == 0;
        // a candidate fits when its footprint is covered by the largest shape AND the brick fits the stage
        const bool covered = mine.ex[0] <= c_box_x[kNumBX - 1] && mine.ex[1] <= c_box_y[kNumBY - 1];
        const unsigned fits = __ballot_sync(0xffffffffu, empty || (covered && bx * by * mine.ex[2] <= kStageVox)) & ((1u << kSlabMax) - 1u);
        const int pick = max(__ffs(~fits) - 1, 1) - 1;      // longest run of fitting candidates, at least one sample
        if (lane == pick) {
            SlabHeader h;
            h.ka = ka, h.kb = (ka < kt1) ? cand : ka;
            h.lx = mine.lo[0], h.ly = mine.lo[1], h.lz = mine.lo[2];
            h.ex = empty ? 0 : bx, h.ey = empty ? 0 : by;
            h.ez = empty ? 0 : min(mine.ex[2], kStageVox / (bx * by));   // a single sample too deep: clip (the rest gathers)
            h.shape = iy * kNumBX + ix;
            sm.hdr[s & 7] = h;
        }
        __syncwarp();
    };
    // issue the brick of (already planned) slab s into its stage
    auto issue = [&](int s) {
        const int st = s % kStages;
        const SlabHeader h = sm.hdr[s & 7];
        const uint32_t plane_bytes = static_cast<uint32_t>(h.ex * h.ey) * 64u;
        if (lane == 0) {
            if (h.ez > 0) mbar_arrive_expect_tx(&sm.full[st], plane_bytes * static_cast<uint32_t>(h.ez));
            else mbar_arrive(&sm.full[st]);
        }
        __syncwarp();
        const uint32_t dst0 = stage0 + static_cast<uint32_t>(st) * (kStageVox * 64);
        if (lane < h.ez)
            tma_load_4d(dst0 + static_cast<uint32_t>(lane) * plane_bytes, &maps.m[h.shape], 0, h.lx, h.ly, zvol + h.lz + lane,
                        &sm.full[st]);
    };
    if (warp == kConsWarps) {
        // ================= producer warp =================
        // only the wait for a free stage and the issue sit on the critical path: slab s + 1 is planned right after slab s
        // has been issued, i.e. while its copies are in flight and the consumers are busy
        if (lane < kNumBX * kNumBY) prefetch_tensormap(&maps.m[lane]);
        plan(0, kt0);
        for (int s = 0; sm.hdr[s & 7].ka < kt1; ++s) {
            mbar_wait(&sm.empty[s % kStages], ((s / kStages) & 1) ^ 1);
            issue(s);
            plan(s + 1, sm.hdr[s & 7].kb);
        }
        return;
    }

    const float4* qv = dens_quad + v * (D + 2) * Hq * Wq;
    const int row_y = Wp * 16, row_z = Hp * Wp * 16;          // float strides of the padded feature volume
    uint32_t choff[4];                                        // byte offset of the chunk this lane reads i-th
#pragma unroll
    for (int e = 0; e < 4; ++e) choff[e] = static_cast<uint32_t>((e ^ rq) << 4);

    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    float T = 1.f, depth = 0.f;

    // Software pipeline, one sample ahead: the footprint of sample k + 1 is computed and its density quad requested before
    // the feature work of sample k, so the global-load latency hides behind the 16 LDS.128 + 32 FFMA2 of the current sample.
    const FootScale fs = {0.5f * static_cast<float>(W - 1), 0.5f * static_cast<float>(H - 1), 0.5f * static_cast<float>(D - 1)};
    const float4* qvc = qv + (static_cast<long long>(1 + c) * Hq + 1) * Wq + 1;     // lane c's plane, border folded in
    auto fetch = [&](int kk, Foot& f, float4& d4) -> bool {     // footprint of sample kk + its density quad (lane c: plane z0 + c)
        if (kTuned) {
            f = sample_foot2(r, sm.zs[kk], fs, D, H, W);
            const bool act = f.in && (kk >= r.k0) && (kk < r.k1);
            if (act) d4 = __ldg(qvc + ((f.z0 * Hq + f.y0) * Wq + f.x0));             // 32-bit index inside one volume
            return act;
        }
        f = sample_foot(r, sm.zs[kk], D, H, W);
        const bool act = f.in && (kk >= r.k0) && (kk < r.k1);
        if (act) d4 = __ldg(qv + (static_cast<long long>(f.z0 + 1 + c) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1));
        return act;
    };
    Foot fn;
    float4 d4n = make_float4(0.f, 0.f, 0.f, 0.f);
    bool actn = false;
    if (kPrefetch && kw1 > kw0) actn = fetch(max(kt0, kw0), fn, d4n);

    for (int s = 0;; ++s) {
        const int st = s % kStages;
        mbar_wait(&sm.full[st], (s / kStages) & 1);
        const SlabHeader h = sm.hdr[s & 7];
        const uint32_t brick = stage0 + static_cast<uint32_t>(st) * (kStageVox * 64);
        const int kend = min(h.kb, kw1);
        // tuned (kMode 1 / 3 / 4): two samples per trip, so the one-ahead (footprint, quad) registers alternate instead of being
        // copied; with all 16 loads of a sample in flight (kMode 5, default) one sample per trip is faster again (0.3175 vs 0.3224 ms
        // at cfg-2, 4.344 vs 4.376 ms at cfg-4; four per trip: 0.344)
#pragma unroll(kMode == 5 ? 1 : (kTuned ? 2 : 1))
        for (int k = max(h.ka, kw0); k < kend; ++k) {
            const float z = sm.zs[k];
            if (!kPrefetch) actn = fetch(k, fn, d4n);
            const Foot f = fn;
            const bool act = actn;
            const float4 d4 = d4n;
            if (kPrefetch && k + 1 < kw1) actn = fetch(k + 1, fn, d4n);
            const float w00 = __fmul_rn(f.wx0, f.wy0), w10 = __fmul_rn(f.wx1, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1),
                        w11 = __fmul_rn(f.wx1, f.wy1);
            float part = 0.f;
            if (act) {          // density: lane c owns plane z0 + c (one 16-byte quad = its four x/y corners)
                const float wz = c ? f.wz1 : f.wz0;
                part = __fmul_rn(w00, wz) * d4.x;
                part = fmaf(__fmul_rn(w10, wz), d4.y, part);
                part = fmaf(__fmul_rn(w01, wz), d4.z, part);
                part = fmaf(__fmul_rn(w11, wz), d4.w, part);
            }
            const float sigma = part + __shfl_xor_sync(0xffffffffu, part, 1);
            const float wk = sigma * T;
            if (wk != 0.f) {    // sigma != 0 implies act
                // features: lane c owns the x-corner x0 + c; weights in ATen's order (wx wy) wz, times the sample weight
                const float wy0 = c ? w10 : w00, wy1 = c ? w11 : w01;
                const float cw[4] = {wk * __fmul_rn(wy0, f.wz0), wk * __fmul_rn(wy1, f.wz0), wk * __fmul_rn(wy0, f.wz1),
                                     wk * __fmul_rn(wy1, f.wz1)};        // index = dz * 2 + dy
                const int xb = f.x0 + 1 - h.lx, yb = f.y0 + 1 - h.ly, zb = f.z0 + 1 - h.lz;
                const bool inbox = (static_cast<unsigned>(xb) + 1u < static_cast<unsigned>(h.ex)) &&
                                   (static_cast<unsigned>(yb) + 1u < static_cast<unsigned>(h.ey)) &&
                                   (static_cast<unsigned>(zb) + 1u < static_cast<unsigned>(h.ez));
                if (inbox && kMode == 3) {
                    // one z-plane at a time: its 8 LDS.128 (two corner rows) in flight before the 16 FFMA2
                    const uint32_t a0 = brick + static_cast<uint32_t>(((zb * h.ey + yb) * h.ex + xb + c) << 6);
                    const uint32_t sy = static_cast<uint32_t>(h.ex) << 6, sz = static_cast<uint32_t>(h.ex * h.ey) << 6;
#pragma unroll
                    for (int dz = 0; dz < 2; ++dz) {
                        const uint32_t ap = a0 + (dz ? sz : 0u);
                        float4 v[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[e] = lds128(ap + choff[e]);
                            v[4 + e] = lds128(ap + sy + choff[e]);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            fma4(acc + 4 * e, cw[2 * dz], v[e]);
                            fma4(acc + 4 * e, cw[2 * dz + 1], v[4 + e]);
                        }
                    }
                } else if (inbox && kMode >= 4) {
                    // all 16 LDS.128 of the sample in flight before the 32 FFMA2
                    const uint32_t a0 = brick + static_cast<uint32_t>(((zb * h.ey + yb) * h.ex + xb + c) << 6);
                    const uint32_t sy = static_cast<uint32_t>(h.ex) << 6, sz = static_cast<uint32_t>(h.ex * h.ey) << 6;
                    float4 v[16];
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {
                        const uint32_t ap = a0 + ((cn & 2) ? sz : 0u) + ((cn & 1) ? sy : 0u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[4 * cn + e] = lds128(ap + choff[e]);
                    }
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn)
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw[cn], v[4 * cn + e]);
                } else if (inbox && kTuned) {
                    // one (dz, dy) corner row at a time: its 4 LDS.128 are issued back to back, then the 8 FFMA2 (the compiler's own
                    // schedule kept 2 loads in flight per lane)
                    const uint32_t a0 = brick + static_cast<uint32_t>(((zb * h.ey + yb) * h.ex + xb + c) << 6);
                    const uint32_t sy = static_cast<uint32_t>(h.ex) << 6, sz = static_cast<uint32_t>(h.ex * h.ey) << 6;
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {
                        const uint32_t ap = a0 + ((cn & 2) ? sz : 0u) + ((cn & 1) ? sy : 0u);
                        float4 v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = lds128(ap + choff[e]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw[cn], v[e]);
                    }
                } else if (inbox) {
                    const uint32_t a0 = brick + static_cast<uint32_t>(((zb * h.ey + yb) * h.ex + xb + c) << 6);
                    const uint32_t sy = static_cast<uint32_t>(h.ex) << 6, sz = static_cast<uint32_t>(h.ex * h.ey) << 6;
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {
                        const uint32_t a = a0 + ((cn & 2) ? sz : 0u) + ((cn & 1) ? sy : 0u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw[cn], lds128(a + choff[e]));
                    }
                } else {
                    const float* p = fvol + ((f.z0 + 1) * Hp + (f.y0 + 1)) * row_y + (f.x0 + 1 + c) * 16;
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {
                        const float* pc = p + ((cn & 2) ? row_z : 0) + ((cn & 1) ? row_y : 0);
#pragma unroll
                        for (int e = 0; e < 4; ++e) fma4(acc + 4 * e, cw[cn], ldg128(pc + (choff[e] >> 2)));
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

    // ---- combine the two x-corner halves of the pair, store (acc[4 e + t] is channel 4 (e ^ rq) + t) ----
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        float* o = out_feat + pix * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if ((e >> 1) == c)
                *reinterpret_cast<float4*>(o + (choff[e] >> 2)) = make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
        if (c == 0) {
            out_sil[pix] = 1.f - T;
            if (out_depth) out_depth[pix] = depth;
        }
    }
}

template <int kStages, int kStageVox, int kWarpsY, bool kPrefetch = true, int kMode = 0>
static int tma_launch_cfg(const char* fn, const TmaMaps& maps, const float* feat_pad, const float4* dens_quad,
                          const int* view2vol, const float* cam12, const float* zs, float* out_feat, float* out_sil,
                          float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P, cudaStream_t st) {
    constexpr int bytes = tma_smem_bytes(kStages, kStageVox), kTH = 4 * kWarpsY, threads = 32 * (4 * kWarpsY + 1);
    static_assert((kWarpsY <= 2 ? 2 : 1) * bytes <= 227 * 1024 && kStages <= kMaxStages, "the CTAs of an SM must fit");
    static_assert(kStageVox >= 16 * 16, "a stage must hold one plane of the largest shape");
    if (int e = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(raymarch_fwd_tma_kernel<kStages, kStageVox, kWarpsY, kPrefetch, kMode>), bytes))
        return e;
    const int tiles_x = (S_w + kTW - 1) / kTW, tiles_y = (S_h + kTH - 1) / kTH;
    dim3 grid(tiles_x * tiles_y, N);
    raymarch_fwd_tma_kernel<kStages, kStageVox, kWarpsY, kPrefetch, kMode><<<grid, threads, bytes, st>>>(
        maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, D, H, W, S_h, S_w, P, tiles_x,
        interleave_views(V, D, H, W));
    return check_launch(fn);
}

int raymarch_fwd_tma1_launch(const char* fn, const float* feat_pad, const float4* dens_quad, const int* view2vol,
                             const float* cam12, const float* zs, float* out_feat, float* out_sil, float* out_depth, int N,
                             int V, int D, int H, int W, int S_h, int S_w, int P, cudaStream_t st);      // raymarch_tma1.cu

int raymarch_fwd_tma_launch(const char* fn, const float* feat_pad, const float4* dens_quad, const int* view2vol,
                            const float* cam12, const float* zs, float* out_feat, float* out_sil, float* out_depth, int N,
                            int V, int D, int H, int W, int S_h, int S_w, int P, cudaStream_t st) {
    static const int ring = [] {                // tuning knob (development): ring shape "stages x voxels per stage"
        const char* e = getenv("FORGE_K1T_RING");
        return e ? atoi(e) : 14;
    }();
    if (ring == 7)          // one lane per ray (raymarch_tma1.cu)
        return raymarch_fwd_tma1_launch(fn, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D, H, W, S_h,
                                        S_w, P, st);
    // tensor maps over feat_pad viewed as [V (D+2)] [H+2] [W+2] [16] fp32, one per static in-plane box shape (host-side
    // encode, ~1 us each; passed to the kernel as __grid_constant__ parameters: no device allocation)
    alignas(64) TmaMaps maps;
    const unsigned long long dims[4] = {16ull, static_cast<unsigned long long>(W + 2), static_cast<unsigned long long>(H + 2),
                                        static_cast<unsigned long long>(D + 2) * V};
    const unsigned long long strides[3] = {64ull, 64ull * (W + 2), 64ull * (W + 2) * (H + 2)};
    for (int iy = 0; iy < kNumBY; ++iy)
        for (int ix = 0; ix < kNumBX; ++ix) {
            const unsigned box[4] = {16u, static_cast<unsigned>(kBoxX[ix]), static_cast<unsigned>(kBoxY[iy]), 1u};
            if (int e = encode_tensor_map(fn, &maps.m[iy * kNumBX + ix], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, feat_pad, dims, strides,
                                          box, CU_TENSOR_MAP_SWIZZLE_NONE))
                return e;
        }
#define FORGE_K1T(S, VOX, WY)                                                                                                \
    return tma_launch_cfg<S, VOX, WY>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D, H, \
                                      W, S_h, S_w, P, st)
    if (ring == 8)          // 880-voxel stages (the most two CTAs per SM can hold)
        return tma_launch_cfg<2, 880, 2>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D, H, W,
                                         S_h, S_w, P, st);
    if (ring == 9)          // no software pipelining of the density quad
        return tma_launch_cfg<2, 832, 2, false>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D,
                                                H, W, S_h, S_w, P, st);
    if (ring == 10)         // tuned instruction stream (A/B against the default)
        return tma_launch_cfg<2, 880, 2, true, 1>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V,
                                                     D, H, W, S_h, S_w, P, st);
    if (ring == 13)         // tuned loop with 16 LDS.128 in flight per lane (A/B)
        return tma_launch_cfg<2, 880, 2, true, 4>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V,
                                                  D, H, W, S_h, S_w, P, st);
    if (ring == 14)         // 16 in flight, one sample per trip (A/B)
        return tma_launch_cfg<2, 880, 2, true, 5>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V,
                                                  D, H, W, S_h, S_w, P, st);
    if (ring == 12)         // tuned loop with 8 LDS.128 in flight per lane (A/B)
        return tma_launch_cfg<2, 880, 2, true, 3>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V,
                                                  D, H, W, S_h, S_w, P, st);
    if (ring == 3) FORGE_K1T(3, 552, 2);
    if (ring == 5) FORGE_K1T(2, 1700, 4);         // 16 x 16 pixel tile, one CTA per SM, twice the slab length
    if (ring == 6) FORGE_K1T(3, 1130, 4);
    if (ring == 1) FORGE_K1T(2, 832, 2);
    if (ring == 2) FORGE_K1T(2, 880, 2);          // round-2 first default (untuned instruction stream)
#undef FORGE_K1T
    // default: the largest stages two CTAs per SM can hold (0.347 vs 0.349 ms at cfg-2, 4.68 vs 4.73 ms at cfg-4) with the tuned
    // sample loop (folded un-normalisation, F2I + I2FP floor, 32-bit quad index) and all 16 LDS.128 of a sample in flight before
    // their 32 FFMA2, one sample per trip: 0.348 -> 0.318 ms at cfg-2, 4.68 -> 4.34 ms at cfg-4 (4 / 8 / 16 loads in flight with
    // two samples per trip: 0.332 / 0.328 / 0.322 ms; FORGE_K1T_RING = 10 / 12 / 13, default 14)
    return tma_launch_cfg<2, 880, 2, true, 5>(fn, maps, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D,
                                              H, W, S_h, S_w, P, st);
}

}  // namespace forge
