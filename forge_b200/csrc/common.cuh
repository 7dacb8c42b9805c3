// Shared device helpers and host-side error plumbing for libforge_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/forge_b200.h"

namespace forge {

void set_error(const std::string& msg);
int fail(const char* fn, const std::string& msg);
int check_launch(const char* fn);
// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (function attributes are per device; a
// process may drive several).  Remembers what was set per (kernel, device); returns 0 or a fail() code.
int ensure_dynamic_smem(const char* fn, const void* kernel, size_t bytes);
// Multiprocessor count of the current device (cached per device); 0 + error string on failure.
int current_sm_count(const char* fn);

// NVTX range around every C-ABI entry point (SURVEY 5 "Tracing"): header-only NVTX v3, a no-op pointer check unless a
// tool (nsys / ncu --nvtx) is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define FORGE_RANGE(name) ::forge::NvtxRange forge_nvtx_range_(name)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- index path (bit-exact contract with ATen GridSampler.h:27-36) ---------------------------
// No FMA contraction on this path: every op is an explicitly rounded intrinsic.
__device__ __forceinline__ float unnormalize_ac(float x, int size) {      // align_corners = True
    return __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), static_cast<float>(size - 1));
}
__device__ __forceinline__ float unnormalize_nac(float x, int size) {     // align_corners = False
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(x, 1.f), static_cast<float>(size)), 1.f), 0.5f);
}

struct Tri {            // trilinear footprint of one sample
    int x0, y0, z0;     // base (floor) voxel
    float wx0, wx1, wy0, wy1, wz0, wz1;   // ATen weights: (x0+1 - ix), (ix - x0), ...
    unsigned mask;      // bit (dz*4 + dy*2 + dx) set when that corner lies inside the volume
};

__device__ __forceinline__ Tri make_tri(float ix, float iy, float iz, int D, int H, int W) {
    Tri t;
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    t.x0 = static_cast<int>(fx);
    t.y0 = static_cast<int>(fy);
    t.z0 = static_cast<int>(fz);
    t.wx1 = __fsub_rn(ix, fx);
    t.wx0 = __fsub_rn(__fadd_rn(fx, 1.f), ix);
    t.wy1 = __fsub_rn(iy, fy);
    t.wy0 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
    t.wz1 = __fsub_rn(iz, fz);
    t.wz0 = __fsub_rn(__fadd_rn(fz, 1.f), iz);
    const unsigned mx = (t.x0 >= 0 && t.x0 < W ? 1u : 0u) | (t.x0 + 1 >= 0 && t.x0 + 1 < W ? 2u : 0u);
    const unsigned my = (t.y0 >= 0 && t.y0 < H ? 1u : 0u) | (t.y0 + 1 >= 0 && t.y0 + 1 < H ? 2u : 0u);
    const unsigned mz = (t.z0 >= 0 && t.z0 < D ? 1u : 0u) | (t.z0 + 1 >= 0 && t.z0 + 1 < D ? 2u : 0u);
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const bool in = ((mx >> (c & 1)) & 1u) && ((my >> ((c >> 1) & 1)) & 1u) && ((mz >> (c >> 2)) & 1u);
        m |= (in ? 1u : 0u) << c;
    }
    t.mask = m;
    return t;
}

// weight of corner c (bit0 = dx, bit1 = dy, bit2 = dz), multiplied x*y*z left to right like ATen
__device__ __forceinline__ float tri_weight(const Tri& t, int c) {
    const float wx = (c & 1) ? t.wx1 : t.wx0;
    const float wy = (c & 2) ? t.wy1 : t.wy0;
    const float wz = (c & 4) ? t.wz1 : t.wz0;
    return __fmul_rn(__fmul_rn(wx, wy), wz);
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    // vectorised fire-and-forget reduction (sm_90+): one 16-byte RED instead of four
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace forge
