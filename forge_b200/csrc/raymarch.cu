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
// voxel apart, so their corner reads share 128-byte lines), a 256-thread CTA a 16x8 pixel tile.
// The two lanes split the density planes (dz = c) and combine with one xor-shuffle.  Samples outside
// the exact ray/volume slab are skipped: under zeros padding they contribute exactly 0 and multiply
// the transmittance by exactly 1.
#include "common.cuh"

namespace forge {

constexpr int kRmThreads = 256;
constexpr int kMaxP = 512;

struct Ray {
    float ox, oy, oz, dx, dy, dz;
    int k0, k1;   // conservative sample range that can touch the volume
};

__device__ __forceinline__ void axis_slab(float o, float d, float lim, float& zlo, float& zhi) {
    if (d != 0.f) {
        const float t1 = (-lim - o) / d, t2 = (lim - o) / d;
        zlo = fmaxf(zlo, fminf(t1, t2));
        zhi = fminf(zhi, fmaxf(t1, t2));
    } else if (fabsf(o) >= lim) {
        zlo = 1.f;
        zhi = 0.f;   // empty
    }
}

__device__ __forceinline__ Ray make_ray(const float* __restrict__ cam, int i, int j, const float* zs, int P, int D, int H,
                                        int W) {
    Ray r;
    const float u = static_cast<float>(j) + 0.5f, v = static_cast<float>(i) + 0.5f;
    r.ox = cam[0];
    r.oy = cam[1];
    r.oz = cam[2];
    r.dx = fmaf(cam[3], u, fmaf(cam[4], v, cam[5]));
    r.dy = fmaf(cam[6], u, fmaf(cam[7], v, cam[8]));
    r.dz = fmaf(cam[9], u, fmaf(cam[10], v, cam[11]));
    // a sample can touch the volume only if every unnormalised coordinate lies in (-1, size):
    // |p_axis| < 1 + 2/(size-1).  Solve for z per axis, widen by one sample on each side.
    float zlo = -3.0e38f, zhi = 3.0e38f;
    axis_slab(r.ox, r.dx, 1.f + 2.f / static_cast<float>(W - 1), zlo, zhi);
    axis_slab(r.oy, r.dy, 1.f + 2.f / static_cast<float>(H - 1), zlo, zhi);
    axis_slab(r.oz, r.dz, 1.f + 2.f / static_cast<float>(D - 1), zlo, zhi);
    r.k0 = 0;
    r.k1 = P;
    if (zlo > zhi) {
        r.k1 = 0;
    } else if (P >= 2) {
        const float z0 = zs[0], dzs = (zs[P - 1] - zs[0]) / static_cast<float>(P - 1);
        if (dzs > 0.f) {
            const float a = fminf(fmaxf((zlo - z0) / dzs, -2.f), static_cast<float>(P) + 2.f);
            const float b = fminf(fmaxf((zhi - z0) / dzs, -2.f), static_cast<float>(P) + 2.f);
            r.k0 = max(0, static_cast<int>(floorf(a)) - 1);
            r.k1 = min(P, static_cast<int>(ceilf(b)) + 2);
        }
    }
    return r;
}

// Sample footprint without masks (the packed layouts make every corner addressable).
struct Foot {
    int x0, y0, z0;
    float wx0, wx1, wy0, wy1, wz0, wz1;
    bool in;   // base voxel inside [-1, size-1] on every axis <=> the sample can touch the volume
};

__device__ __forceinline__ Foot sample_foot(const Ray& r, float z, int D, int H, int W) {
    // points = origins + lengths * directions, rounded like the reference's separate mul and add;
    // un-normalisation on the bit-exact index path (common.cuh)
    const float ix = unnormalize_ac(__fadd_rn(r.ox, __fmul_rn(z, r.dx)), W);
    const float iy = unnormalize_ac(__fadd_rn(r.oy, __fmul_rn(z, r.dy)), H);
    const float iz = unnormalize_ac(__fadd_rn(r.oz, __fmul_rn(z, r.dz)), D);
    Foot f;
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    f.x0 = static_cast<int>(fx);
    f.y0 = static_cast<int>(fy);
    f.z0 = static_cast<int>(fz);
    f.wx1 = __fsub_rn(ix, fx);
    f.wx0 = __fsub_rn(__fadd_rn(fx, 1.f), ix);
    f.wy1 = __fsub_rn(iy, fy);
    f.wy0 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
    f.wz1 = __fsub_rn(iz, fz);
    f.wz0 = __fsub_rn(__fadd_rn(fz, 1.f), iz);
    f.in = (static_cast<unsigned>(f.x0 + 1) <= static_cast<unsigned>(W)) &&
           (static_cast<unsigned>(f.y0 + 1) <= static_cast<unsigned>(H)) &&
           (static_cast<unsigned>(f.z0 + 1) <= static_cast<unsigned>(D));
    return f;
}

struct f8 {
    float v[8];
};

__device__ __forceinline__ f8 ldg256(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                   "=f"(r.v[7])
                 : "l"(p));
    return r;
}

__global__ void __launch_bounds__(kRmThreads, 2)
raymarch_fwd_kernel(const float* __restrict__ feat_pad, const float4* __restrict__ dens_quad,
                    const int* __restrict__ view2vol, const float* __restrict__ cam12, const float* __restrict__ zs_g,
                    float* __restrict__ out_feat, float* __restrict__ out_sil, float* __restrict__ out_depth, int D,
                    int H, int W, int Sh, int Sw, int P, int tiles_x) {
    __shared__ float zs[kMaxP];
    __shared__ float cam[12];
    const int n = blockIdx.y;
    for (int k = threadIdx.x; k < P; k += kRmThreads) zs[k] = zs_g[k];
    if (threadIdx.x < 12) cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    __syncthreads();

    // 16x8 pixel tile per CTA, 4x4 patch per warp, 2 lanes per ray
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane & 1, q = lane >> 1;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int j = tx * 16 + (warp & 3) * 4 + (q & 3);
    const int i = ty * 8 + (warp >> 2) * 4 + (q >> 2);
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

    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
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
            float fs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) fs[e] = 0.f;
#pragma unroll
            for (int cn = 0; cn < 8; ++cn) {
                const float wxy = (cn & 2) ? ((cn & 1) ? w11 : w01) : ((cn & 1) ? w10 : w00);
                const float w = __fmul_rn(wxy, (cn & 4) ? f.wz1 : f.wz0);
                const f8 val = ldg256(p + ((cn & 4) ? row_z : 0) + ((cn & 2) ? row_y : 0) + ((cn & 1) ? 16 : 0));
#pragma unroll
                for (int e = 0; e < 8; ++e) fs[e] = fmaf(val.v[e], w, fs[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(wk, fs[e], acc[e]);
            depth = fmaf(wk, z, depth);
        }
        T = T * (1.f - sigma);
    }
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        float4* o = reinterpret_cast<float4*>(out_feat + pix * 16 + c * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
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
// Mapping: 4 lanes per ray (lane c owns channels 4c..4c+3), 8x8 pixel tile per CTA.  Features come
// from the padded volume (and grad_feat goes to a padded volume of the same shape); densities are
// read from / scattered to the plain [V][D][H][W] layout, 2 corners per lane.
constexpr int kRaysPerCta = kRmThreads / 4;

__device__ __forceinline__ Tri sample_tri(const Ray& r, float z, int D, int H, int W) {
    const float px = __fadd_rn(r.ox, __fmul_rn(z, r.dx));
    const float py = __fadd_rn(r.oy, __fmul_rn(z, r.dy));
    const float pz = __fadd_rn(r.oz, __fmul_rn(z, r.dz));
    return make_tri(unnormalize_ac(px, W), unnormalize_ac(py, H), unnormalize_ac(pz, D), D, H, W);
}

// density at the sample: lane c of the quad fetches corners (dz = c>>1, dy = c&1, dx = 0/1)
__device__ __forceinline__ float quad_density(const Tri& t, const float* __restrict__ dens, int H, int W, int c) {
    const int c0 = ((c >> 1) << 2) | ((c & 1) << 1);
    const int zz = t.z0 + (c >> 1), yy = t.y0 + (c & 1);
    const long long row = (static_cast<long long>(zz) * H + yy) * W + t.x0;
    float part = 0.f;
    if ((t.mask >> c0) & 1u) part = __ldg(dens + row) * tri_weight(t, c0);
    if ((t.mask >> (c0 + 1)) & 1u) part = fmaf(__ldg(dens + row + 1), tri_weight(t, c0 + 1), part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    return part;
}

__global__ void __launch_bounds__(kRmThreads)
raymarch_bwd_kernel(const float4* __restrict__ feat_pad, const float* __restrict__ dens,
                    const int* __restrict__ view2vol, const float* __restrict__ cam12, const float* __restrict__ zs_g,
                    const float4* __restrict__ g_feat, const float* __restrict__ g_sil,
                    const float* __restrict__ g_depth, float* __restrict__ grad_feat_pad,
                    float* __restrict__ grad_dens, float* __restrict__ grad_cam, int D, int H, int W, int Sh, int Sw,
                    int P, int tiles_x) {
    extern __shared__ float smem[];
    float* zs = smem;                       // [P]
    float* cam = zs + P;                    // [12]
    float* red = cam + 12;                  // [12][8 warps]
    float* s_sig = red + 12 * (kRmThreads / 32);   // [P][64]
    float* s_a = s_sig + P * kRaysPerCta;
    float* s_T = s_a + P * kRaysPerCta;

    const int n = blockIdx.y;
    for (int k = threadIdx.x; k < P; k += kRmThreads) zs[k] = zs_g[k];
    if (threadIdx.x < 12) cam[threadIdx.x] = cam12[n * 12 + threadIdx.x];
    __syncthreads();

    const bool need_feat = grad_feat_pad != nullptr, need_dens = grad_dens != nullptr, need_cam = grad_cam != nullptr;
    const int c = threadIdx.x & 3, ray = threadIdx.x >> 2;
    const int warp = threadIdx.x >> 5, q = (threadIdx.x & 31) >> 2;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int j = tx * 8 + (warp & 1) * 4 + (q & 3);
    const int i = ty * 8 + (warp >> 1) * 2 + (q >> 2);
    const bool valid = (i < Sh) && (j < Sw);
    Ray r = make_ray(cam, i, j, zs, P, D, H, W);
    if (!valid) r.k1 = 0;
    int kw0 = r.k1 > r.k0 ? r.k0 : P, kw1 = r.k1 > r.k0 ? r.k1 : 0;
#pragma unroll
    for (int s = 16; s >= 4; s >>= 1) {
        kw0 = min(kw0, __shfl_xor_sync(0xffffffffu, kw0, s));
        kw1 = max(kw1, __shfl_xor_sync(0xffffffffu, kw1, s));
    }

    const int Wp = W + 2, Hp = H + 2;
    const long long vol = static_cast<long long>(view2vol[n]);
    const long long volp = vol * (D + 2) * Hp * Wp;
    const float4* fv = feat_pad + volp * 4 + c;       // 4 float4 per padded voxel
    const float* dv = dens + vol * D * H * W;
    float* gfv = need_feat ? grad_feat_pad + volp * 16 + 4 * c : nullptr;
    float* gdv = need_dens ? grad_dens + vol * D * H * W : nullptr;

    float4 gF = make_float4(0.f, 0.f, 0.f, 0.f);
    float gO = 0.f, gD = 0.f;
    if (valid) {
        const long long pix = (static_cast<long long>(n) * Sh + i) * Sw + j;
        gF = g_feat[pix * 4 + c];
        gO = g_sil[pix];
        if (g_depth) gD = g_depth[pix];
    }
    const float sx = 0.5f * static_cast<float>(W - 1), sy = 0.5f * static_cast<float>(H - 1),
                sz = 0.5f * static_cast<float>(D - 1);
    float go0 = 0.f, go1 = 0.f, go2 = 0.f, gd0 = 0.f, gd1 = 0.f, gd2 = 0.f;   // per-lane partials

    // ---- pass A: front to back ----
    float T = 1.f;
    for (int k = kw0; k < kw1; ++k) {
        const bool act = (k >= r.k0) && (k < r.k1);
        const float z = zs[k];
        Tri t = sample_tri(r, z, D, H, W);
        if (!act) t.mask = 0;
        const float sigma = quad_density(t, dv, H, W, c);
        const float wk = sigma * T;
        float a_part = 0.f, gix = 0.f, giy = 0.f, giz = 0.f;
        if (t.mask) {
            const bool scatter = need_feat && (wk != 0.f);
            // mask != 0 => base voxel in [-1, size-1]: all 8 padded corners are addressable
            const int base = ((t.z0 + 1) * Hp + (t.y0 + 1)) * Wp + (t.x0 + 1);
#pragma unroll
            for (int cn = 0; cn < 8; ++cn) {
                if ((t.mask >> cn) & 1u) {
                    const int vox = base + ((cn & 4) ? Hp * Wp : 0) + ((cn & 2) ? Wp : 0) + (cn & 1);
                    const float4 val = __ldg(fv + static_cast<long long>(vox) * 4);
                    const float q4 = fmaf(gF.x, val.x, fmaf(gF.y, val.y, fmaf(gF.z, val.z, gF.w * val.w)));
                    const float wx = (cn & 1) ? t.wx1 : t.wx0, wy = (cn & 2) ? t.wy1 : t.wy0, wz = (cn & 4) ? t.wz1 : t.wz0;
                    const float wyz = wy * wz;
                    const float w = wx * wyz;
                    a_part = fmaf(w, q4, a_part);
                    gix = fmaf((cn & 1) ? wyz : -wyz, q4, gix);
                    giy = fmaf((cn & 2) ? wx * wz : -(wx * wz), q4, giy);
                    giz = fmaf((cn & 4) ? wx * wy : -(wx * wy), q4, giz);
                    if (scatter) {
                        const float cw = wk * w;
                        red_add_v4(gfv + static_cast<long long>(vox) * 16,
                                   make_float4(cw * gF.x, cw * gF.y, cw * gF.z, cw * gF.w));
                    }
                }
            }
        }
        a_part += __shfl_xor_sync(0xffffffffu, a_part, 1);
        a_part += __shfl_xor_sync(0xffffffffu, a_part, 2);
        if (act) {
            if (c == 0) {
                s_sig[k * kRaysPerCta + ray] = sigma;
                s_a[k * kRaysPerCta + ray] = fmaf(gD, z, a_part);
                s_T[k * kRaysPerCta + ray] = T;
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
    __syncwarp();

    // ---- pass B: back to front (per-quad loop, no shuffles inside) ----
    if (need_dens || need_cam) {
        float Bk = -gO;
        const int c0 = ((c >> 1) << 2) | ((c & 1) << 1);
        for (int k = r.k1 - 1; k >= r.k0; --k) {
            const float sigma = s_sig[k * kRaysPerCta + ray], a = s_a[k * kRaysPerCta + ray], Tk = s_T[k * kRaysPerCta + ray];
            const float dsig = Tk * (a - Bk);
            Bk = fmaf(a, sigma, (1.f - sigma) * Bk);
            const float z = zs[k];
            const Tri t = sample_tri(r, z, D, H, W);
            const long long row = (static_cast<long long>(t.z0 + (c >> 1)) * H + (t.y0 + (c & 1))) * W + t.x0;
            const float wy = (c & 1) ? t.wy1 : t.wy0, wz = (c >> 1) ? t.wz1 : t.wz0;
            float gix = 0.f, giy = 0.f, giz = 0.f;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                if ((t.mask >> (c0 + dx)) & 1u) {
                    const float wx = dx ? t.wx1 : t.wx0;
                    if (need_dens) atomicAdd(gdv + row + dx, dsig * (wx * wy * wz));
                    if (need_cam) {
                        const float val = __ldg(dv + row + dx);
                        gix = fmaf(dx ? wy * wz : -(wy * wz), val, gix);
                        giy = fmaf((c & 1) ? wx * wz : -(wx * wz), val, giy);
                        giz = fmaf((c >> 1) ? wx * wy : -(wx * wy), val, giz);
                    }
                }
            }
            const float px = dsig * gix * sx, py = dsig * giy * sy, pz = dsig * giz * sz;
            go0 += px;
            go1 += py;
            go2 += pz;
            gd0 = fmaf(z, px, gd0);
            gd1 = fmaf(z, py, gd1);
            gd2 = fmaf(z, pz, gd2);
        }
    }
    __syncwarp();

    // ---- camera gradient: 12 floats per view, reduced over the CTA ----
    if (need_cam) {
        const float u = static_cast<float>(j) + 0.5f, v = static_cast<float>(i) + 0.5f;
        float g12[12] = {go0, go1, go2, gd0 * u, gd0 * v, gd0, gd1 * u, gd1 * v, gd1, gd2 * u, gd2 * v, gd2};
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            float x = valid ? g12[e] : 0.f;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
            if (lane == 0) red[e * (kRmThreads / 32) + warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < kRmThreads / 32; ++w) x += red[threadIdx.x * (kRmThreads / 32) + w];
            atomicAdd(grad_cam + n * 12 + threadIdx.x, x);
        }
    }
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

extern "C" int forge_raymarch_fwd(const float* feat_pad, const float* dens_quad, const int* view2vol,
                                  const float* cam12, const float* zs, float* out_feat, float* out_sil,
                                  float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                                  void* stream) {
    using namespace forge;
    const char* fn = "forge_raymarch_fwd";
    if (!feat_pad || !dens_quad || !view2vol || !cam12 || !zs || !out_feat || !out_sil) return fail(fn, "null pointer");
    if (int e = raymarch_check(fn, N, V, D, H, W, S_h, S_w, P)) return e;
    if (P > kMaxP) return fail(fn, "n_pts_per_ray exceeds 512");
    if ((reinterpret_cast<uintptr_t>(feat_pad) & 31u) || !aligned16(dens_quad) || !aligned16(out_feat))
        return fail(fn, "feat_pad must be 32-byte aligned, dens_quad / out_feat 16-byte aligned");
    const int tiles_x = (S_w + 15) / 16, tiles_y = (S_h + 7) / 8;
    dim3 grid(tiles_x * tiles_y, N);
    raymarch_fwd_kernel<<<grid, kRmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        feat_pad, reinterpret_cast<const float4*>(dens_quad), view2vol, cam12, zs, out_feat, out_sil, out_depth, D, H, W,
        S_h, S_w, P, tiles_x);
    return check_launch(fn);
}

extern "C" int forge_raymarch_bwd(const float* feat_pad, const float* dens, const int* view2vol, const float* cam12,
                                  const float* zs, const float* g_feat, const float* g_sil, const float* g_depth,
                                  float* grad_feat_pad, float* grad_dens, float* grad_cam12, int N, int V, int D, int H,
                                  int W, int S_h, int S_w, int P, void* stream) {
    using namespace forge;
    const char* fn = "forge_raymarch_bwd";
    if (!feat_pad || !dens || !view2vol || !cam12 || !zs || !g_feat || !g_sil) return fail(fn, "null pointer");
    if (!grad_feat_pad && !grad_dens && !grad_cam12) return 0;
    if (int e = raymarch_check(fn, N, V, D, H, W, S_h, S_w, P)) return e;
    if (!aligned16(feat_pad) || !aligned16(g_feat) || (grad_feat_pad && !aligned16(grad_feat_pad)))
        return fail(fn, "feat_pad / g_feat / grad_feat_pad must be 16-byte aligned");
    const size_t smem = sizeof(float) * (static_cast<size_t>(P) + 12 + 12 * (kRmThreads / 32) +
                                         3 * static_cast<size_t>(P) * kRaysPerCta);
    if (smem > 227 * 1024) return fail(fn, "n_pts_per_ray too large for the backward pass (max 300)");
    static thread_local size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(raymarch_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return fail(fn, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
        smem_set = smem;
    }
    const int tiles_x = (S_w + 7) / 8, tiles_y = (S_h + 7) / 8;
    dim3 grid(tiles_x * tiles_y, N);
    raymarch_bwd_kernel<<<grid, kRmThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(feat_pad), dens, view2vol, cam12, zs, reinterpret_cast<const float4*>(g_feat),
        g_sil, g_depth, grad_feat_pad, grad_dens, grad_cam12, D, H, W, S_h, S_w, P, tiles_x);
    return check_launch(fn);
}
