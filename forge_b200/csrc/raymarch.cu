// K1: fused camera -> ray -> trilinear fetch -> emission-absorption composite (forward + backward).
//
// Replaces reference models/volume_render.py:53-63 (PyTorch3D NDCGridRaysampler + VolumeSampler +
// EmissionAbsorptionRaymarcher + README.md:26-33 depth patch).  Nothing between the packed volume
// and the [N,S,S,16]+sil+depth images touches HBM.
//
// Inputs are the PACKED render volumes built by forge_pack_volume (layout.cu):
//   feat_pad  [V][D+2][H+2][W+2][16]  channels-last with a one-voxel zero border, so the 8 trilinear
//             corners of any sample that can touch the volume are valid addresses and zeros padding
//             needs no per-corner predicate;
//   dens_quad [V][D+2][H+1][W+1][4]   = (d(z,y,x), d(z,y,x+1), d(z,y+1,x), d(z,y+1,x+1)), zero outside:
//             one aligned 16-byte load returns the four density corners of a z-plane.
//
// Forward mapping: 2 lanes per ray, lane c owns feature channels 8c..8c+7 and fetches each corner
// with one 256-bit load (LDG.E.256); a warp marches a 4x4 pixel patch (neighbouring rays are ~0.5
// voxel apart, so their corner reads share 128-byte lines), a 128-thread CTA a 16x4 pixel tile.
// The two lanes split the density planes (dz = c) and combine with one xor-shuffle.  Samples outside
// the exact ray/volume slab are skipped: under zeros padding they contribute exactly 0 and multiply
// the transmittance by exactly 1.
#include <algorithm>
#include <cstdlib>

#include "raymarch_common.cuh"

namespace forge {

int raymarch_fwd_tma_launch(const char* fn, const float* feat_pad, const float4* dens_quad, const int* view2vol,
                            const float* cam12, const float* zs, float* out_feat, float* out_sil, float* out_depth, int N,
                            int V, int D, int H, int W, int S_h, int S_w, int P, cudaStream_t st);      // raymarch_tma.cu

// CTA = kWX x kWY warps, each marching a 4x4 pixel patch: (4 kWX) x (4 kWY) pixel tile
template <int kMinBlocks, int kWX, int kWY>
__global__ void __launch_bounds__(32 * kWX * kWY, kMinBlocks)
raymarch_fwd_kernel(const float* __restrict__ feat_pad, const float4* __restrict__ dens_quad,
                    const int* __restrict__ view2vol, const float* __restrict__ cam12, const float* __restrict__ zs_g,
                    float* __restrict__ out_feat, float* __restrict__ out_sil, float* __restrict__ out_depth, int D,
                    int H, int W, int Sh, int Sw, int P, int tiles_x, int interleave) {
    __shared__ float zs[kMaxP];
    __shared__ float cam[12];
    // Heavy-first schedule: CTAs are dispatched in (blockIdx.x, then blockIdx.y) order; map that order to
    // (tile rank, view) with the views interleaved and the tiles ranked centre-out, so the long CTAs (rays through the
    // middle of the volume) start first and the last wave consists of the short border tiles (ncu: SMs were idle
    // 9 % of the launch with the row-major order).
    const int order = blockIdx.y * gridDim.x + blockIdx.x, n_views = gridDim.y, tiles_y = gridDim.x / tiles_x;
    // views are interleaved only while all packed volumes fit in L2 together (host decision); otherwise view by view
    const int rank = interleave ? order / n_views : static_cast<int>(blockIdx.x);
    const int n = interleave ? order - rank * n_views : static_cast<int>(blockIdx.y);
    const int ri = rank / tiles_x, ci = rank - ri * tiles_x;
    const int ty = centre_out(ri, tiles_y), tx = centre_out(ci, tiles_x);
    for (int k = threadIdx.x; k < P; k += 32 * kWX * kWY) zs[k] = zs_g[k];
    if (threadIdx.x < 12) cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    __syncthreads();

    // 4x4 patch per warp, 2 lanes per ray
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane & 1, q = lane >> 1;
    const int j = tx * (4 * kWX) + (warp % kWX) * 4 + (q & 3);
    const int i = ty * (4 * kWY) + (warp / kWX) * 4 + (q >> 2);
    const bool valid = (i < Sh) && (j < Sw);
    Ray r = make_ray(cam, i, j, zs, P, D, H, W);
    if (!valid) r.k1 = 0;
    // warp-uniform loop bounds (the pair shuffle needs every lane of the warp in the loop)
    int kw0 = r.k1 > r.k0 ? r.k0 : P, kw1 = r.k1 > r.k0 ? r.k1 : 0;
#pragma unroll
    for (int s = 16; s >= 2; s >>= 1) {
        kw0 = min(kw0, __shfl_xor_sync(0xffffffffu, kw0, s));
        kw1 = max(kw1, __shfl_xor_sync(0xffffffffu, kw1, s));
    }

    const int Wp = W + 2, Hp = H + 2, Wq = W + 1, Hq = H + 1;
    const int v = view2vol[n];
    const float* fv = feat_pad + static_cast<long long>(v) * (D + 2) * Hp * Wp * 16 + c * 8;
    const float4* qv = dens_quad + static_cast<long long>(v) * (D + 2) * Hq * Wq;
    const int row_y = Wp * 16, row_z = Hp * Wp * 16;   // float strides of the padded feature volume

    float2 acc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = make_float2(0.f, 0.f);
    float T = 1.f, depth = 0.f;
    for (int k = kw0; k < kw1; ++k) {
        const float z = zs[k];
        const Foot f = sample_foot(r, z, D, H, W);
        const bool act = f.in && (k >= r.k0) && (k < r.k1);
        // ATen weight order: (wx * wy) * wz
        const float w00 = __fmul_rn(f.wx0, f.wy0), w10 = __fmul_rn(f.wx1, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1),
                    w11 = __fmul_rn(f.wx1, f.wy1);
        float part = 0.f;
        if (act) {
            const float wz = c ? f.wz1 : f.wz0;
            const float4 d4 = __ldg(qv + (static_cast<long long>(f.z0 + 1 + c) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1));
            part = __fmul_rn(w00, wz) * d4.x;
            part = fmaf(__fmul_rn(w10, wz), d4.y, part);
            part = fmaf(__fmul_rn(w01, wz), d4.z, part);
            part = fmaf(__fmul_rn(w11, wz), d4.w, part);
        }
        const float sigma = part + __shfl_xor_sync(0xffffffffu, part, 1);
        const float wk = sigma * T;
        if (wk != 0.f) {   // sigma != 0 implies act
            const float* p = fv + ((f.z0 + 1) * Hp + (f.y0 + 1)) * row_y + (f.x0 + 1) * 16;
            // packed FFMA2 (two fp32 FMAs per instruction on sm_100, per-component rounding = fmaf)
            float2 fs[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) fs[e] = make_float2(0.f, 0.f);
#pragma unroll
            for (int cn = 0; cn < 8; ++cn) {
                const float wxy = (cn & 2) ? ((cn & 1) ? w11 : w01) : ((cn & 1) ? w10 : w00);
                const float w = __fmul_rn(wxy, (cn & 4) ? f.wz1 : f.wz0);
                const f8 val = ldg256(p + ((cn & 4) ? row_z : 0) + ((cn & 2) ? row_y : 0) + ((cn & 1) ? 16 : 0));
                const float2 w2 = make_float2(w, w);
#pragma unroll
                for (int e = 0; e < 4; ++e) fs[e] = __ffma2_rn(make_float2(val.v[2 * e], val.v[2 * e + 1]), w2, fs[e]);
            }
            const float2 wk2 = make_float2(wk, wk);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = __ffma2_rn(wk2, fs[e], acc[e]);
            depth = fmaf(wk, z, depth);
        }
        T = T * (1.f - sigma);
    }
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        float4* o = reinterpret_cast<float4*>(out_feat + pix * 16 + c * 8);
        o[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
        o[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
        if (c == 0) {
            out_sil[pix] = 1.f - T;
            if (out_depth) out_depth[pix] = depth;
        }
    }
}

// ---- backward ------------------------------------------------------------------------------------
// With a_k = g_F . f_k + g_D z_k, T_k = prod_{j<k}(1 - s_j) and the suffix recurrence
//   B_{P-1} = -g_O,   B_{k-1} = a_k s_k + (1 - s_k) B_k
// the density gradient is dL/ds_k = T_k (a_k - B_k): no division by (1 - s_k), so s_k = 1 and
// s_k > 1 are handled exactly like torch.cumprod's backward.  Pass A marches front-to-back
// (re-fetching the features, scattering grad_feat, stashing s_k, a_k, T_k in shared memory),
// pass B walks back-to-front (dL/ds_k, grad_dens scatter, d s/d p for the camera gradient).
// d L / d p_k is accumulated per lane into d L / d o and d L / d dir and reduced once per CTA into
// the 12 camera floats of the view.
//
// Same mapping and packed inputs as the forward kernel (2 lanes per ray, 256-bit corner loads, density
// quads, 16x4 pixel tile per CTA).  grad_feat goes to a volume in feat_pad layout, grad_dens to a
// zero-bordered [V][D+2][H+2][W+2] volume, so neither scatter needs bounds predicates.
constexpr int kBwThreads = 128;                 // 4 warps = 16x4 pixel tile (finer CTAs balance better than 16x8)
constexpr int kRaysPerCta = kBwThreads / 2;

// kMerge: rays of a warp whose samples share the base voxel (0.5-voxel ray spacing: 16 rays of a 4x4 patch fall into ~6.7 distinct
// base voxels) are summed by one leader pair before the RED: REDs cost ~1.2 cycles per active lane in the LSU whatever their
// width, and that issue rate -- not L2 -- bounded the unmerged kernel (20 RED warp-instructions x 32 lanes per 16 ray-samples).
template <bool kMerge, int kMinBlocks>
__global__ void __launch_bounds__(kBwThreads, kMinBlocks)
raymarch_bwd_kernel(const float* __restrict__ feat_pad, const float4* __restrict__ dens_quad,
                    const int* __restrict__ view2vol, const float* __restrict__ cam12, const float* __restrict__ zs_g,
                    const float* __restrict__ g_feat, const float* __restrict__ g_sil,
                    const float* __restrict__ g_depth, float* __restrict__ grad_feat_pad,
                    float* __restrict__ grad_dens_pad, float* __restrict__ grad_cam, float* __restrict__ workspace,
                    float4* __restrict__ grad_quad, int D, int H, int W, int Sh, int Sw, int P, int tiles_x, int interleave) {
    __shared__ float zs[kMaxP];
    __shared__ float cam[12];
    __shared__ float red[12 * (kBwThreads / 32)];
    __shared__ __align__(16) float gFs[kMerge ? kRaysPerCta * 16 : 4];   // upstream feature gradient per ray, in scatter order
    __shared__ __align__(16) float cws[kMerge ? kRaysPerCta * 8 : 4];    // this sample's 8 corner weights (x w_k) per ray
    // per-CTA stash [P][3][64 rays] in the caller's workspace (L2-resident between the two passes);
    // keeping it out of shared memory leaves the whole unified L1 to the corner gathers
    float* stash = workspace + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * P * 3 * kRaysPerCta;
    // heavy-first schedule, as in the forward kernel: views interleaved, tiles ranked centre-out
    const int order = blockIdx.y * gridDim.x + blockIdx.x, n_views = gridDim.y, tiles_y = gridDim.x / tiles_x;
    const int rank = interleave ? order / n_views : static_cast<int>(blockIdx.x);
    const int n = interleave ? order - rank * n_views : static_cast<int>(blockIdx.y);
    const int ty = centre_out(rank / tiles_x, tiles_y), tx = centre_out(rank % tiles_x, tiles_x);

    for (int k = threadIdx.x; k < P; k += kBwThreads) zs[k] = zs_g[k];
    if (threadIdx.x < 12) cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    __syncthreads();

    const bool need_feat = grad_feat_pad != nullptr, need_dens = grad_dens_pad != nullptr, need_cam = grad_cam != nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane & 1, q = lane >> 1, ray = threadIdx.x >> 1;
    const int j = tx * 16 + (warp & 3) * 4 + (q & 3);
    const int i = ty * 4 + (q >> 2);
    const bool valid = (i < Sh) && (j < Sw);
    Ray r = make_ray(cam, i, j, zs, P, D, H, W);
    if (!valid) r.k1 = 0;
    int kw0 = r.k1 > r.k0 ? r.k0 : P, kw1 = r.k1 > r.k0 ? r.k1 : 0;
#pragma unroll
    for (int s = 16; s >= 2; s >>= 1) {
        kw0 = min(kw0, __shfl_xor_sync(0xffffffffu, kw0, s));
        kw1 = max(kw1, __shfl_xor_sync(0xffffffffu, kw1, s));
    }

    const int Wp = W + 2, Hp = H + 2, Wq = W + 1, Hq = H + 1;
    const long long v = view2vol[n];
    const long long volp = v * (D + 2) * Hp * Wp;
    const float* fv = feat_pad + volp * 16 + c * 8;
    const float4* qv = dens_quad + v * (D + 2) * Hq * Wq;
    float* gdv = need_dens ? grad_dens_pad + volp : nullptr;
    float4* gqv = (need_dens && grad_quad) ? grad_quad + v * (D + 2) * Hq * Wq : nullptr;    // dens_quad layout (kernel-side scratch)
    const int row_y = Wp * 16, row_z = Hp * Wp * 16;

    float gF[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) gF[e] = 0.f;
    float gO = 0.f, gD = 0.f;
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        const float4* gp = reinterpret_cast<const float4*>(g_feat + pix * 16 + c * 8);
        const float4 g0 = gp[0], g1 = gp[1];
        gF[0] = g0.x, gF[1] = g0.y, gF[2] = g0.z, gF[3] = g0.w;
        gF[4] = g1.x, gF[5] = g1.y, gF[6] = g1.z, gF[7] = g1.w;
        gO = g_sil[pix];
        if (g_depth) gD = g_depth[pix];
    }
    const float sx = 0.5f * static_cast<float>(W - 1), sy = 0.5f * static_cast<float>(H - 1),
                sz = 0.5f * static_cast<float>(D - 1);
    float go0 = 0.f, go1 = 0.f, go2 = 0.f, gd0 = 0.f, gd1 = 0.f, gd2 = 0.f;   // per-lane partials
    // scatter layout: each RED.128 instruction of a ray should fill whole 32-byte sectors, so lane c
    // adds channels [4c, 4c+4) (first sector) and [8+4c, 8+4c+4) (second sector); half of those
    // upstream-gradient values live in the partner lane
    float sA[4], sB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float mine_lo = gF[e], mine_hi = gF[4 + e];
        const float other_lo = __shfl_xor_sync(0xffffffffu, mine_lo, 1), other_hi = __shfl_xor_sync(0xffffffffu, mine_hi, 1);
        sA[e] = c ? other_hi : mine_lo;     // channels 4c+e of the ray:      c=0 -> own gF[e],   c=1 -> lane0's gF[4+e]
        sB[e] = c ? mine_hi : other_lo;     // channels 8+4c+e of the ray:    c=0 -> lane1's gF[e], c=1 -> own gF[4+e]
    }
    float* gfs = need_feat ? grad_feat_pad + volp * 16 + c * 4 : nullptr;
    if (kMerge && need_feat) {
        float4* g = reinterpret_cast<float4*>(&gFs[(ray * 2 + c) * 8]);
        g[0] = make_float4(sA[0], sA[1], sA[2], sA[3]);
        g[1] = make_float4(sB[0], sB[1], sB[2], sB[3]);
        __syncwarp();                   // read back by lanes of the same warp only
    }

    // ---- pass A: front to back ----
    float T = 1.f;
    for (int k = kw0; k < kw1; ++k) {
        const float z = zs[k];
        const Foot f = sample_foot(r, z, D, H, W);
        const bool act = f.in && (k >= r.k0) && (k < r.k1);
        const float w00 = __fmul_rn(f.wx0, f.wy0), w10 = __fmul_rn(f.wx1, f.wy0), w01 = __fmul_rn(f.wx0, f.wy1),
                    w11 = __fmul_rn(f.wx1, f.wy1);
        float part = 0.f;
        if (act) {
            const float wz = c ? f.wz1 : f.wz0;
            const float4 d4 = __ldg(qv + (static_cast<long long>(f.z0 + 1 + c) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1));
            part = __fmul_rn(w00, wz) * d4.x;
            part = fmaf(__fmul_rn(w10, wz), d4.y, part);
            part = fmaf(__fmul_rn(w01, wz), d4.z, part);
            part = fmaf(__fmul_rn(w11, wz), d4.w, part);
        }
        const float sigma = part + __shfl_xor_sync(0xffffffffu, part, 1);
        const float wk = sigma * T;
        float a_part = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
        float cwv[4] = {0.f, 0.f, 0.f, 0.f};         // kMerge: w_k x weight of the corners of plane dz = c
        int off = 0;
        if (act) {
            const bool scatter = need_feat && (wk != 0.f);
            off = ((f.z0 + 1) * Hp + (f.y0 + 1)) * row_y + (f.x0 + 1) * 16;
#pragma unroll
            for (int cn = 0; cn < 8; ++cn) {
                const int o = off + ((cn & 4) ? row_z : 0) + ((cn & 2) ? row_y : 0) + ((cn & 1) ? 16 : 0);
                const f8 val = ldg256(fv + o);
                // g_F . f over 8 channels as packed FFMA2 on channel pairs
                float2 q2 = __ffma2_rn(make_float2(gF[0], gF[1]), make_float2(val.v[0], val.v[1]), make_float2(0.f, 0.f));
#pragma unroll
                for (int e = 1; e < 4; ++e) q2 = __ffma2_rn(make_float2(gF[2 * e], gF[2 * e + 1]), make_float2(val.v[2 * e], val.v[2 * e + 1]), q2);
                const float qd = q2.x + q2.y;
                const float wx = (cn & 1) ? f.wx1 : f.wx0, wy = (cn & 2) ? f.wy1 : f.wy0, wz = (cn & 4) ? f.wz1 : f.wz0;
                const float wxy = (cn & 2) ? ((cn & 1) ? w11 : w01) : ((cn & 1) ? w10 : w00);
                const float w = wxy * wz;
                a_part = fmaf(w, qd, a_part);
                gix = fmaf((cn & 1) ? wy * wz : -(wy * wz), qd, gix);
                giy = fmaf((cn & 2) ? wx * wz : -(wx * wz), qd, giy);
                giz = fmaf((cn & 4) ? wxy : -wxy, qd, giz);
                if (kMerge) {
                    if ((cn >> 2) == c) cwv[cn & 3] = wk * w;
                } else if (scatter) {
                    const float cw = wk * w;
                    red_add_v4(gfs + o, make_float4(cw * sA[0], cw * sA[1], cw * sA[2], cw * sA[3]));
                    red_add_v4(gfs + o + 8, make_float4(cw * sB[0], cw * sB[1], cw * sB[2], cw * sB[3]));
                }
            }
        }
        if (kMerge && need_feat) {          // warp-uniform: every lane takes part in the match
            const bool scat = act && (wk != 0.f);
            __syncwarp();                   // the previous sample's reads of cws are done
            if (scat) *reinterpret_cast<float4*>(&cws[ray * 8 + 4 * c]) = make_float4(cwv[0], cwv[1], cwv[2], cwv[3]);
            const unsigned grp = __match_any_sync(0xffffffffu, scat ? off : ~lane);      // lanes (ray pairs) with the same base voxel
            __syncwarp();
            // The first two member rays of a group take one z-plane of corners each (a lone ray takes both, one after the other), so
            // the warp spends ~group size iterations instead of twice that with the rest of the group idle.
            const unsigned members = grp & 0x55555555u;                  // one bit per member ray
            const int gsize = __popc(members), rank = __popc(members & ((1u << (lane & 30)) - 1u));
            auto reduce_plane = [&](int hz) {                            // sum the group's contributions to the 4 corners of plane dz = hz
                float2 a[4][2], b[4][2];                                 // [corner][channel pair]: lane c owns channels 4c.. and 8 + 4c..
#pragma unroll
                for (int e = 0; e < 4; ++e) a[e][0] = a[e][1] = b[e][0] = b[e][1] = make_float2(0.f, 0.f);
                for (unsigned rr = members; rr; rr &= rr - 1) {
                    const int mray = (warp * 32 + __ffs(rr) - 1) >> 1;
                    const float4 w4 = *reinterpret_cast<const float4*>(&cws[mray * 8 + 4 * hz]);
                    const float4 ga = *reinterpret_cast<const float4*>(&gFs[(mray * 2 + c) * 8]);
                    const float4 gb = *reinterpret_cast<const float4*>(&gFs[(mray * 2 + c) * 8 + 4]);
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {                        // packed FFMA2: two channels per instruction
                        const float2 w2 = make_float2(wv[e], wv[e]);
                        a[e][0] = __ffma2_rn(w2, make_float2(ga.x, ga.y), a[e][0]);
                        a[e][1] = __ffma2_rn(w2, make_float2(ga.z, ga.w), a[e][1]);
                        b[e][0] = __ffma2_rn(w2, make_float2(gb.x, gb.y), b[e][0]);
                        b[e][1] = __ffma2_rn(w2, make_float2(gb.z, gb.w), b[e][1]);
                    }
                }
                const int oz = off + (hz ? row_z : 0);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int o = oz + ((e & 2) ? row_y : 0) + ((e & 1) ? 16 : 0);
                    red_add_v4(gfs + o, make_float4(a[e][0].x, a[e][0].y, a[e][1].x, a[e][1].y));
                    red_add_v4(gfs + o + 8, make_float4(b[e][0].x, b[e][0].y, b[e][1].x, b[e][1].y));
                }
            };
            if (scat && rank < 2) reduce_plane(rank);
            if (scat && gsize == 1) reduce_plane(1);
        }
        a_part += __shfl_xor_sync(0xffffffffu, a_part, 1);
        if (act) {
            if (c == 0) {
                float* st = stash + static_cast<size_t>(k) * 3 * kRaysPerCta + ray;
                st[0] = sigma;
                st[kRaysPerCta] = fmaf(gD, z, a_part);
                st[2 * kRaysPerCta] = T;
            }
            const float px = wk * gix * sx, py = wk * giy * sy, pz = wk * giz * sz;
            go0 += px;
            go1 += py;
            go2 += pz;
            gd0 = fmaf(z, px, gd0);
            gd1 = fmaf(z, py, gd1);
            gd2 = fmaf(z, pz, gd2);
        }
        T = T * (1.f - sigma);
    }
    __syncwarp();     // lane 0's stash writes are visible to lane 1 of the pair (same warp, global memory)

    // ---- pass B: back to front (per-ray loop, no shuffles inside; lane c owns density plane dz = c) ----
    if (need_dens || need_cam) {
        float Bk = -gO;
        for (int k = r.k1 - 1; k >= r.k0; --k) {
            const float z = zs[k];
            const Foot f = sample_foot(r, z, D, H, W);
            if (!f.in) continue;            // never stashed: sigma = 0, B unchanged
            const float* st = stash + static_cast<size_t>(k) * 3 * kRaysPerCta + ray;
            const float sigma = st[0], a = st[kRaysPerCta], Tk = st[2 * kRaysPerCta];
            const float dsig = Tk * (a - Bk);
            Bk = fmaf(a, sigma, (1.f - sigma) * Bk);
            const float wz = c ? f.wz1 : f.wz0;
            if (need_dens && gqv) {         // one 16-byte RED per lane instead of four scalar ones (REDs cost per lane, not per byte)
                const float dz = dsig * wz;
                red_add_v4(reinterpret_cast<float*>(gqv + (static_cast<long long>(f.z0 + 1 + c) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1)),
                           make_float4(dz * (f.wx0 * f.wy0), dz * (f.wx1 * f.wy0), dz * (f.wx0 * f.wy1), dz * (f.wx1 * f.wy1)));
            } else if (need_dens) {
                float* g = gdv + (static_cast<long long>(f.z0 + 1 + c) * Hp + (f.y0 + 1)) * Wp + (f.x0 + 1);
                const float dz = dsig * wz;
                atomicAdd(g, dz * (f.wx0 * f.wy0));
                atomicAdd(g + 1, dz * (f.wx1 * f.wy0));
                atomicAdd(g + Wp, dz * (f.wx0 * f.wy1));
                atomicAdd(g + Wp + 1, dz * (f.wx1 * f.wy1));
            }
            if (need_cam) {
                const float4 d4 = __ldg(qv + (static_cast<long long>(f.z0 + 1 + c) * Hq + (f.y0 + 1)) * Wq + (f.x0 + 1));
                // d sigma / d(ix, iy, iz) restricted to this lane's z-plane
                const float gx_ = wz * (f.wy0 * (d4.y - d4.x) + f.wy1 * (d4.w - d4.z));
                const float gy_ = wz * (f.wx0 * (d4.z - d4.x) + f.wx1 * (d4.w - d4.y));
                const float pl = f.wx0 * f.wy0 * d4.x + f.wx1 * f.wy0 * d4.y + f.wx0 * f.wy1 * d4.z + f.wx1 * f.wy1 * d4.w;
                const float gz_ = c ? pl : -pl;
                const float px = dsig * gx_ * sx, py = dsig * gy_ * sy, pz = dsig * gz_ * sz;
                go0 += px;
                go1 += py;
                go2 += pz;
                gd0 = fmaf(z, px, gd0);
                gd1 = fmaf(z, py, gd1);
                gd2 = fmaf(z, pz, gd2);
            }
        }
    }
    __syncwarp();

    // ---- camera gradient: 12 floats per view, reduced over the CTA ----
    if (need_cam) {
        const float u = static_cast<float>(j) + 0.5f, vv = static_cast<float>(i) + 0.5f;
        float g12[12] = {go0, go1, go2, gd0 * u, gd0 * vv, gd0, gd1 * u, gd1 * vv, gd1, gd2 * u, gd2 * vv, gd2};
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            float x = valid ? g12[e] : 0.f;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
            if (lane == 0) red[e * (kBwThreads / 32) + warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < kBwThreads / 32; ++w) x += red[threadIdx.x * (kBwThreads / 32) + w];
            atomicAdd(grad_cam + n * 12 + threadIdx.x, x);
        }
    }
}

// grad_dens_pad[zp][Y][X] += the four quad components that alias padded voxel (zp, Y, X): quad (zp, yq, xq) holds the
// gradients of the padded voxels (yq, xq), (yq, xq+1), (yq+1, xq), (yq+1, xq+1), yq in [0, H], xq in [0, W].
__global__ void __launch_bounds__(256)
fold_dens_quads_kernel(const float4* __restrict__ quad, float* __restrict__ grad_dens_pad, long long planes, int H, int W) {
    const int Hp = H + 2, Wp = W + 2, Hq = H + 1, Wq = W + 1;
    const long long total = planes * Hp * Wp;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
        const int X = static_cast<int>(i % Wp), Y = static_cast<int>((i / Wp) % Hp);
        const float4* q = quad + (i / (static_cast<long long>(Wp) * Hp)) * Hq * Wq;
        float g = 0.f;
        if (Y < Hq && X < Wq) g += q[Y * Wq + X].x;
        if (Y < Hq && X >= 1) g += q[Y * Wq + X - 1].y;
        if (Y >= 1 && X < Wq) g += q[(Y - 1) * Wq + X].z;
        if (Y >= 1 && X >= 1) g += q[(Y - 1) * Wq + X - 1].w;
        grad_dens_pad[i] += g;
    }
}

static long long bwd_stash_bytes(int N, int S_h, int S_w, int P) {
    const long long tiles = static_cast<long long>((S_w + 15) / 16) * ((S_h + 3) / 4);
    return tiles * N * P * 3 * kRaysPerCta * static_cast<long long>(sizeof(float));
}
static long long bwd_quad_bytes(int V, int D, int H, int W) {
    return static_cast<long long>(V) * (D + 2) * (H + 1) * (W + 1) * 16;
}

static int raymarch_check(const char* fn, int N, int V, int D, int H, int W, int S_h, int S_w, int P) {
    if (N <= 0 || V <= 0 || S_h <= 0 || S_w <= 0 || P <= 0) return fail(fn, "non-positive size");
    if (D < 2 || H < 2 || W < 2) return fail(fn, "volume sides must be >= 2");
    if (N > 65535) return fail(fn, "more than 65535 views in one launch");
    if (static_cast<long long>(D + 2) * (H + 2) * (W + 2) * 16 >= 2147483647LL)
        return fail(fn, "volume too large (padded feature volume must stay below 2^31 floats)");
    return 0;
}

}  // namespace forge

namespace forge {

static int raymarch_fwd_args(const char* fn, const float* feat_pad, const float* dens_quad, const int* view2vol,
                             const float* cam12, const float* zs, float* out_feat, float* out_sil, int N, int V, int D, int H,
                             int W, int S_h, int S_w, int P) {
    if (!feat_pad || !dens_quad || !view2vol || !cam12 || !zs || !out_feat || !out_sil) return fail(fn, "null pointer");
    if (int e = raymarch_check(fn, N, V, D, H, W, S_h, S_w, P)) return e;
    if (P > kMaxP) return fail(fn, "n_pts_per_ray exceeds 512");
    if ((reinterpret_cast<uintptr_t>(feat_pad) & 31u) || !aligned16(dens_quad) || !aligned16(out_feat))
        return fail(fn, "feat_pad must be 32-byte aligned, dens_quad / out_feat 16-byte aligned");
    return 0;
}

static int raymarch_fwd_gather_launch(const char* fn, const float* feat_pad, const float* dens_quad, const int* view2vol,
                                      const float* cam12, const float* zs, float* out_feat, float* out_sil,
                                      float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                      void* stream) {
    static const int min_blocks = [] {          // tuning knob (development): resident CTAs per SM the kernel is built for
        const char* e = getenv("FORGE_K1_MINBLOCKS");
        return e ? atoi(e) : 3;         // measured on B200, cfg-2: 2 -> 0.411 ms, 3 -> 0.395 ms, 4 -> 0.394 ms
    }();
    static const int shape = [] {               // tuning knob (development): CTA shape in warps, "XY" (41 = 4 x 1 warps)
        const char* e = getenv("FORGE_K1_SHAPE");
        return e ? atoi(e) : 41;
    }();
    const int wx = shape / 10, wy = shape % 10;
    const int tiles_x = (S_w + 4 * wx - 1) / (4 * wx), tiles_y = (S_h + 4 * wy - 1) / (4 * wy);
    dim3 grid(tiles_x * tiles_y, N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4* dq = reinterpret_cast<const float4*>(dens_quad);
    const int interleave = interleave_views(V, D, H, W);
    // min_blocks counts 256-thread equivalents: the kernels are built for min_blocks * 8 resident warps per SM
#define FORGE_K1_LAUNCH(WX, WY)                                                                                            \
    do {                                                                                                                   \
        constexpr int kW = WX * WY;                                                                                        \
        if (min_blocks == 3)                                                                                               \
            raymarch_fwd_kernel<24 / kW, WX, WY><<<grid, 32 * kW, 0, st>>>(feat_pad, dq, view2vol, cam12, zs, out_feat, out_sil, \
                                                                          out_depth, D, H, W, S_h, S_w, P, tiles_x, interleave); \
        else if (min_blocks == 4)                                                                                          \
            raymarch_fwd_kernel<32 / kW, WX, WY><<<grid, 32 * kW, 0, st>>>(feat_pad, dq, view2vol, cam12, zs, out_feat, out_sil, \
                                                                          out_depth, D, H, W, S_h, S_w, P, tiles_x, interleave); \
        else                                                                                                               \
            raymarch_fwd_kernel<16 / kW, WX, WY><<<grid, 32 * kW, 0, st>>>(feat_pad, dq, view2vol, cam12, zs, out_feat, out_sil, \
                                                                          out_depth, D, H, W, S_h, S_w, P, tiles_x, interleave); \
    } while (0)
    if (shape == 42) FORGE_K1_LAUNCH(4, 2);
    else if (shape == 22) FORGE_K1_LAUNCH(2, 2);
    else if (shape == 21) FORGE_K1_LAUNCH(2, 1);
    else if (shape == 11) FORGE_K1_LAUNCH(1, 1);
    else FORGE_K1_LAUNCH(4, 1);
#undef FORGE_K1_LAUNCH
    return check_launch(fn);
}

}  // namespace forge

extern "C" int forge_raymarch_fwd_gather(const float* feat_pad, const float* dens_quad, const int* view2vol,
                                         const float* cam12, const float* zs, float* out_feat, float* out_sil,
                                         float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                         void* stream) {
    FORGE_RANGE("forge_raymarch_fwd_gather");
    using namespace forge;
    const char* fn = "forge_raymarch_fwd_gather";
    if (int e = raymarch_fwd_args(fn, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, N, V, D, H, W, S_h, S_w, P))
        return e;
    return raymarch_fwd_gather_launch(fn, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D, H, W,
                                      S_h, S_w, P, stream);
}

extern "C" int forge_raymarch_fwd_tma(const float* feat_pad, const float* dens_quad, const int* view2vol,
                                      const float* cam12, const float* zs, float* out_feat, float* out_sil,
                                      float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                      void* stream) {
    FORGE_RANGE("forge_raymarch_fwd_tma");
    using namespace forge;
    const char* fn = "forge_raymarch_fwd_tma";
    if (int e = raymarch_fwd_args(fn, feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, N, V, D, H, W, S_h, S_w, P))
        return e;
    return raymarch_fwd_tma_launch(fn, feat_pad, reinterpret_cast<const float4*>(dens_quad), view2vol, cam12, zs, out_feat,
                                   out_sil, out_depth, N, V, D, H, W, S_h, S_w, P, static_cast<cudaStream_t>(stream));
}

extern "C" int forge_raymarch_fwd(const float* feat_pad, const float* dens_quad, const int* view2vol,
                                  const float* cam12, const float* zs, float* out_feat, float* out_sil,
                                  float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                  void* stream) {
    if (int e = forge::raymarch_fwd_args("forge_raymarch_fwd", feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, N, V,
                                         D, H, W, S_h, S_w, P))
        return e;
    // default formulation: TMA-staged bricks (measured on B200: 0.350 vs 0.367 ms at cfg-2, 4.73 vs 6.09 ms at cfg-4, equal at
    // cfg-1); FORGE_K1_IMPL=gather|tma overrides (read once)
    static const int impl_tma = [] {
        const char* e = getenv("FORGE_K1_IMPL");
        return e ? (e[0] == 't') : 1;
    }();
    return impl_tma ? forge_raymarch_fwd_tma(feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D, H,
                                             W, S_h, S_w, P, stream)
                    : forge_raymarch_fwd_gather(feat_pad, dens_quad, view2vol, cam12, zs, out_feat, out_sil, out_depth, N, V, D,
                                                H, W, S_h, S_w, P, stream);
}

extern "C" long long forge_raymarch_bwd_workspace(int N, int V, int D, int H, int W, int S_h, int S_w, int P) {
    using namespace forge;
    if (N <= 0 || V <= 0 || D <= 0 || H <= 0 || W <= 0 || S_h <= 0 || S_w <= 0 || P <= 0) return 0;
    return bwd_stash_bytes(N, S_h, S_w, P) + bwd_quad_bytes(V, D, H, W);
}

extern "C" int forge_raymarch_bwd(const float* feat_pad, const float* dens_quad, const int* view2vol,
                                  const float* cam12, const float* zs, const float* g_feat, const float* g_sil,
                                  const float* g_depth, float* grad_feat_pad, float* grad_dens_pad, float* grad_cam12,
                                  float* workspace, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                  void* stream) {
    FORGE_RANGE("forge_raymarch_bwd");
    using namespace forge;
    const char* fn = "forge_raymarch_bwd";
    if (!feat_pad || !dens_quad || !view2vol || !cam12 || !zs || !g_feat || !g_sil) return fail(fn, "null pointer");
    if (!grad_feat_pad && !grad_dens_pad && !grad_cam12) return 0;
    if (!workspace) return fail(fn, "null workspace (size it with forge_raymarch_bwd_workspace)");
    if (int e = raymarch_check(fn, N, V, D, H, W, S_h, S_w, P)) return e;
    if (P > kMaxP) return fail(fn, "n_pts_per_ray exceeds 512");
    if ((reinterpret_cast<uintptr_t>(feat_pad) & 31u) || !aligned16(dens_quad) || !aligned16(g_feat) ||
        (grad_feat_pad && !aligned16(grad_feat_pad)))
        return fail(fn, "feat_pad must be 32-byte aligned, dens_quad / g_feat / grad_feat_pad 16-byte aligned");
    const int tiles_x = (S_w + 15) / 16, tiles_y = (S_h + 3) / 4;
    dim3 grid(tiles_x * tiles_y, N);
    static const int merge = [] {          // tuning knob (development): bit 0 = same-base merge of the feature REDs,
        const char* e = getenv("FORGE_K1B_MERGE");      // bit 1 = density gradient through quads; 0 = the round-1 scatter
        return e ? atoi(e) : 3;
    }();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // density gradient: accumulated as quads in the workspace (behind the stash), folded into grad_dens_pad afterwards
    float4* quad = nullptr;
    if (grad_dens_pad && (merge & 2)) {
        quad = reinterpret_cast<float4*>(reinterpret_cast<char*>(workspace) + bwd_stash_bytes(N, S_h, S_w, P));
        if (!aligned16(quad)) return fail(fn, "workspace must be 16-byte aligned");
        if (cudaMemsetAsync(quad, 0, static_cast<size_t>(bwd_quad_bytes(V, D, H, W)), st) != cudaSuccess) return check_launch(fn);
    }
    const float4* dq = reinterpret_cast<const float4*>(dens_quad);
    static const int occ_env = [] {         // tuning knob (development): resident CTAs per SM the kernel is compiled for
        const char* e = getenv("FORGE_K1B_OCC");
        return e ? atoi(e) : 0;
    }();
#define FORGE_K1B(MERGE, OCC)                                                                                                      \
    raymarch_bwd_kernel<MERGE, OCC><<<grid, kBwThreads, 0, st>>>(feat_pad, dq, view2vol, cam12, zs, g_feat, g_sil, g_depth, grad_feat_pad, \
                                                                 grad_dens_pad, grad_cam12, workspace, quad, D, H, W, S_h, S_w, P, tiles_x, \
                                                                 interleave_views(V, D, H, W))
    // measured on B200 (cfg-2, 4 / 5 / 6 CTAs per SM = 128 / 96 / 80 registers): all gradients 1.181 / 1.128 / 1.207 ms (the merged
    // kernel spills at 80), pose only 0.688 / 0.678 / 0.661 ms
    if ((merge & 1) && grad_feat_pad) {
        const int occ = occ_env ? occ_env : 5;
        if (occ == 4) FORGE_K1B(true, 4);
        else if (occ == 6) FORGE_K1B(true, 6);
        else FORGE_K1B(true, 5);
    } else {
        const int occ = occ_env ? occ_env : 6;
        if (occ == 4) FORGE_K1B(false, 4);
        else if (occ == 5) FORGE_K1B(false, 5);
        else FORGE_K1B(false, 6);
    }
#undef FORGE_K1B
    if (quad) {
        const long long planes = static_cast<long long>(V) * (D + 2), total = planes * (H + 2) * (W + 2);
        const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
        fold_dens_quads_kernel<<<blocks, 256, 0, st>>>(quad, grad_dens_pad, planes, H, W);
    }
    return check_launch(fn);
}
