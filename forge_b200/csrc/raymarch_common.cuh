// Device pieces shared by the raymarch kernels (raymarch.cu: direct L1 gathers; raymarch_tma.cu: TMA-staged bricks).
#pragma once
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

// Same footprint with a shorter instruction stream (K1-T's tuned loop): `h` = 0.5 (size - 1) per axis precomputed,
//   ((x + 1) * 0.5) * (size - 1) == (x + 1) * (0.5 (size - 1))   bit for bit: the multiplication by 0.5 is exact (x + 1 is
//   never subnormal: it is 0 or a multiple of 2^-25 for |x| near 1), so both round the same real number once;
//   floor through F2I + I2FP (one conversion-unit instruction instead of two: FRND + F2I), equal wherever the sample
//   can be `in` (|i| < 2^31).
struct FootScale {
    float hx, hy, hz;
};
__device__ __forceinline__ Foot sample_foot2(const Ray& r, float z, const FootScale& s, int D, int H, int W) {
    const float ix = __fmul_rn(__fadd_rn(__fadd_rn(r.ox, __fmul_rn(z, r.dx)), 1.f), s.hx);
    const float iy = __fmul_rn(__fadd_rn(__fadd_rn(r.oy, __fmul_rn(z, r.dy)), 1.f), s.hy);
    const float iz = __fmul_rn(__fadd_rn(__fadd_rn(r.oz, __fmul_rn(z, r.dz)), 1.f), s.hz);
    Foot f;
    f.x0 = __float2int_rd(ix);
    f.y0 = __float2int_rd(iy);
    f.z0 = __float2int_rd(iz);
    const float fx = static_cast<float>(f.x0), fy = static_cast<float>(f.y0), fz = static_cast<float>(f.z0);
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

// i-th element of the sequence c, c+1, c-1, c+2, c-2, ... over [0, n) with c the (lower) middle
__device__ __forceinline__ int centre_out(int i, int n) {
    const int c = (n - 1) >> 1;
    return (i & 1) ? c + ((i + 1) >> 1) : c - (i >> 1);
}


// Interleave the views in the dispatch order only while all packed feature volumes fit in L2 together (cfg-2: 74 MB);
// beyond that (cfg-4: 1.1 GB) a view-by-view order keeps one or two volumes hot at a time.
inline int interleave_views(int V, int D, int H, int W) {
    const long long bytes = static_cast<long long>(V) * (D + 2) * (H + 2) * (W + 2) * 64;
    return bytes <= 96LL * 1024 * 1024;
}

}  // namespace forge
