// Pieces shared by the TMA-staged raymarch kernels (raymarch_tma.cu: 2 lanes per ray; raymarch_tma1.cu: 1 lane per ray).
#pragma once
#include "raymarch_common.cuh"
#include "tensormap.cuh"

namespace forge {

using namespace async_;

constexpr int kTW = 16;                           // pixel tile of a CTA: 16 x (4 kWarpsY); 4 kWarpsY consumer warps + 1 producer
constexpr int kSlabMax = 8;                       // most samples per slab
constexpr int kMaxStages = 4;

struct SlabHeader {
    int ka, kb;          // sample range [ka, kb)
    int lx, ly, lz;      // box origin in padded voxel coordinates
    int ex, ey, ez;      // box extent = pitches of the brick (0 = nothing resident)
    int shape;           // tensor-map index
};

struct TmaSmem {
    unsigned long long full[kMaxStages], empty[kMaxStages];
    SlabHeader hdr[8];                   // slab s lives in hdr[s & 7] (planned up to kStages + 1 slabs ahead)
    float cam[12];
    int kt0, kt1;
    float zs[kMaxP];
};
constexpr int tma_smem_bytes(int stages, int stage_vox) { return stages * stage_vox * 64 + static_cast<int>(sizeof(TmaSmem)); }

struct Box {
    int lo[3], ex[3];
    __device__ __forceinline__ int vol() const { return ex[0] * ex[1] * ex[2]; }
};

// Box (padded voxel coordinates, clipped to the padded volume) of all corner footprints of the samples k in [ka, kb) of
// the tile's rays: positions are multilinear in (pixel, depth), so the 4 corner rays at the 2 end depths bound them.
__device__ __forceinline__ Box slab_box(const float* cam, const float* zs, int ka, int kb, float u0, float u1, float v0,
                                        float v1, int D, int H, int W) {
    const float za = zs[ka], zb = zs[kb - 1];
    const int size[3] = {W, H, D};
    Box b;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float m0 = cam[3 + 3 * a], m1 = cam[4 + 3 * a], m2 = cam[5 + 3 * a], o = cam[a];
        float lo = 3.0e38f, hi = -3.0e38f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float d = fmaf(m0, (c & 1) ? u1 : u0, fmaf(m1, (c & 2) ? v1 : v0, m2));
            const float pa = fmaf(za, d, o), pb = fmaf(zb, d, o);
            lo = fminf(lo, fminf(pa, pb));
            hi = fmaxf(hi, fmaxf(pa, pb));
        }
        // padded voxel coordinate = (p + 1) / 2 * (size - 1) + 1; base = floor, upper corner = base + 1
        const float s = 0.5f * static_cast<float>(size[a] - 1);
        lo = fmaxf(fminf((lo + 1.f) * s + 1.f - 1e-3f, 3.0e4f), -3.0e4f);
        hi = fmaxf(fminf((hi + 1.f) * s + 1.f + 1e-3f, 3.0e4f), -3.0e4f);
        int l = static_cast<int>(floorf(lo)), h = static_cast<int>(floorf(hi)) + 1;
        l = max(l, 0);
        h = min(h, size[a] + 1);
        b.lo[a] = l;
        b.ex[a] = max(h - l + 1, 0);
    }
    if (b.ex[0] == 0 || b.ex[1] == 0 || b.ex[2] == 0) b.ex[0] = b.ex[1] = b.ex[2] = 0;
    return b;
}

__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma4(float* acc, float w, const float4 v) {
    const float2 w2 = make_float2(w, w);
    float2 a = __ffma2_rn(make_float2(v.x, v.y), w2, make_float2(acc[0], acc[1]));
    float2 b = __ffma2_rn(make_float2(v.z, v.w), w2, make_float2(acc[2], acc[3]));
    acc[0] = a.x, acc[1] = a.y, acc[2] = b.x, acc[3] = b.y;
}


}  // namespace forge
