// Elementwise stages of the backward pass of the tensor-core ConvGRU cell (ops._GruTC, constant weights: the gradient flows
// to the views and the initial state only).  Between the two transposed convolutions of a step (forge_conv3d_tc, plain mode)
// the chain rule through h' = h (1 - u) + c u, c = tanh(.), u, r = sigmoid(.) and h r is three streaming passes over dense
// channels-last rows [N = B D H W][C] -- every operand touched once, float4-vectorised, HBM-bound:
//
//   stage A  (dh', u, c, h)                 -> d_o = dh' u (1 - c^2)  [bf16, operand of the transposed out-gate convolution],
//                                              dgu = dh' (c - h) u (1 - u),  dh_dir = dh' (1 - u)
//   stage B  (g1 = [dx_o | d(h r)], r, h, dgu, dh_dir)
//                                           -> dg = [dgu | d(hr) h r (1 - r)]  [bf16, operand of the transposed gate convolution],
//                                              dh_acc = dh_dir + d(hr) r
//   stage C  (g1, g2 = [dx_g | dh_g], dh_acc) -> dx = dx_o + dx_g,  dh = dh_acc + dh_g
#include <cuda_bf16.h>

#include "common.cuh"

namespace forge {
namespace gtb {

constexpr int kThreads = 256;

__device__ __forceinline__ uint2 pack4(float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
#define FORGE_V4(expr_x, expr_y, expr_z, expr_w) make_float4(expr_x, expr_y, expr_z, expr_w)

// C4 = C / 4; index e runs over N * C4 float4 groups
__global__ void __launch_bounds__(kThreads) stage_a(const float4* __restrict__ dh, const float4* __restrict__ u,
                                                     const float4* __restrict__ c, const float4* __restrict__ h,
                                                     uint2* __restrict__ d_o, float4* __restrict__ dgu,
                                                     float4* __restrict__ dh_dir, long long n4) {
    for (long long e = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; e < n4; e += static_cast<long long>(gridDim.x) * kThreads) {
        const float4 g = __ldg(dh + e), uu = __ldg(u + e), cc = __ldg(c + e), hh = __ldg(h + e);
#define A1(f) (g.f * uu.f * (1.f - cc.f * cc.f))
#define A2(f) (g.f * (cc.f - hh.f) * uu.f * (1.f - uu.f))
#define A3(f) (g.f * (1.f - uu.f))
        d_o[e] = pack4(FORGE_V4(A1(x), A1(y), A1(z), A1(w)));
        dgu[e] = FORGE_V4(A2(x), A2(y), A2(z), A2(w));
        dh_dir[e] = FORGE_V4(A3(x), A3(y), A3(z), A3(w));
#undef A1
#undef A2
#undef A3
    }
}

__global__ void __launch_bounds__(kThreads) stage_b(const float4* __restrict__ g1, const float4* __restrict__ r,
                                                     const float4* __restrict__ h, const float4* __restrict__ dgu,
                                                     const float4* __restrict__ dh_dir, uint2* __restrict__ dg,
                                                     float4* __restrict__ dh_acc, long long n4, int C4) {
    for (long long e = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; e < n4; e += static_cast<long long>(gridDim.x) * kThreads) {
        const long long row = e / C4;
        const int cq = static_cast<int>(e - row * C4);
        const float4 dhr = __ldg(g1 + row * 2 * C4 + C4 + cq), rr = __ldg(r + e), hh = __ldg(h + e), gu = __ldg(dgu + e),
                     dd = __ldg(dh_dir + e);
#define B1(f) (dhr.f * hh.f * rr.f * (1.f - rr.f))
#define B2(f) (dd.f + dhr.f * rr.f)
        dg[row * 2 * C4 + cq] = pack4(gu);
        dg[row * 2 * C4 + C4 + cq] = pack4(FORGE_V4(B1(x), B1(y), B1(z), B1(w)));
        dh_acc[e] = FORGE_V4(B2(x), B2(y), B2(z), B2(w));
#undef B1
#undef B2
    }
}

__global__ void __launch_bounds__(kThreads) stage_c(const float4* __restrict__ g1, const float4* __restrict__ g2,
                                                     const float4* __restrict__ dh_acc, float4* __restrict__ dx,
                                                     float4* __restrict__ dh, long long n4, int C4) {
    for (long long e = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; e < n4; e += static_cast<long long>(gridDim.x) * kThreads) {
        const long long row = e / C4;
        const int cq = static_cast<int>(e - row * C4);
        const float4 a = __ldg(g1 + row * 2 * C4 + cq), b = __ldg(g2 + row * 2 * C4 + cq), hg = __ldg(g2 + row * 2 * C4 + C4 + cq),
                     acc = __ldg(dh_acc + e);
        dx[e] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        dh[e] = make_float4(acc.x + hg.x, acc.y + hg.y, acc.z + hg.z, acc.w + hg.w);
    }
}

static int grid_for(long long n4, int sms) {
    const long long want = (n4 + kThreads - 1) / kThreads;
    const long long cap = static_cast<long long>(sms) * 8;
    return static_cast<int>(want < cap ? want : cap);
}

}  // namespace gtb
}  // namespace forge

extern "C" int forge_gru_tc_bwd(int stage, const float* a0, const float* a1, const float* a2, const float* a3,
                                const float* a4, void* o_bf16, float* o0, float* o1, long long N, int C, void* stream) {
    FORGE_RANGE("forge_gru_tc_bwd");
    using namespace forge;
    using namespace forge::gtb;
    const char* fn = "forge_gru_tc_bwd";
    if (N <= 0 || C <= 0 || C % 4) return fail(fn, "C must be a positive multiple of 4");
    const int sms = current_sm_count(fn);
    if (sms <= 0) return 1;
    const long long n4 = N * (C / 4);
    const int C4 = C / 4, grid = grid_for(n4, sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define F4(p) reinterpret_cast<const float4*>(p)
    if (stage == 0) {           // a0 dh', a1 u, a2 c, a3 h -> o_bf16 d_o, o0 dgu, o1 dh_dir
        if (!a0 || !a1 || !a2 || !a3 || !o_bf16 || !o0 || !o1) return fail(fn, "stage A: null pointer");
        stage_a<<<grid, kThreads, 0, st>>>(F4(a0), F4(a1), F4(a2), F4(a3), static_cast<uint2*>(o_bf16), reinterpret_cast<float4*>(o0),
                                           reinterpret_cast<float4*>(o1), n4);
    } else if (stage == 1) {    // a0 g1 [N,2C], a1 r, a2 h, a3 dgu, a4 dh_dir -> o_bf16 dg [N,2C], o0 dh_acc
        if (!a0 || !a1 || !a2 || !a3 || !a4 || !o_bf16 || !o0) return fail(fn, "stage B: null pointer");
        stage_b<<<grid, kThreads, 0, st>>>(F4(a0), F4(a1), F4(a2), F4(a3), F4(a4), static_cast<uint2*>(o_bf16),
                                           reinterpret_cast<float4*>(o0), n4, C4);
    } else if (stage == 2) {    // a0 g1 [N,2C], a1 g2 [N,2C], a2 dh_acc -> o0 dx, o1 dh
        if (!a0 || !a1 || !a2 || !o0 || !o1) return fail(fn, "stage C: null pointer");
        stage_c<<<grid, kThreads, 0, st>>>(F4(a0), F4(a1), F4(a2), reinterpret_cast<float4*>(o0), reinterpret_cast<float4*>(o1), n4, C4);
    } else {
        return fail(fn, "stage must be 0, 1 or 2");
    }
#undef F4
    return check_launch(fn);
}
