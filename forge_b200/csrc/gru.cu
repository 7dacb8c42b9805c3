// ConvGRU cell, elementwise stages (reference models/fusion.py:21-35).  The two 3-D convolutions of a cell stay cuDNN;
// everything between them -- sigmoid, split, h * r, cat, tanh, 1 - u, the lerp -- is ~13 elementwise launches per view in
// the reference (and ~25 in its backward pass), each streaming 17-67 MB.  Here it is two kernels forward, two backward:
//
//   gate:  xhr = cat(x, h * sigmoid(g[:, C:]))                           g = conv_gate(cat(x, h))   [B, 2C, S]
//   out:   h'  = h * (1 - u) + tanh(o) * u,  u = sigmoid(g[:, :C])       o = out_gate(xhr)          [B,  C, S]
//
// Tensors are dense NCDHW ("ncs") or channels-last ("nsc", what K2 emits and the tensor-core convs want); h and x may
// carry their own batch stride (x_t is a slice of the [B, t, ...] view sequence).  The conv outputs g / o (and their
// gradients) may be fp32 or bf16 (autocast); h, x, xhr and h' are fp32.  HBM-bound: every operand is touched once.
#include <cuda_bf16.h>

#include <initializer_list>

#include "common.cuh"

namespace forge {
namespace gru {

constexpr int kThreads = 256;

struct View {          // element (b, c, s) of a [B, CC, S] tensor seen through channel offset c0
    long long bs;      // batch stride in elements
    int CC, c0;
};

// offset of element (b, c, s); cl: channels-last ([B][S][CC]) or channel-major ([B][CC][S])
__device__ __forceinline__ long long at(const View& v, int cl, int S, int b, int c, int s) {
    return cl ? b * v.bs + static_cast<long long>(s) * v.CC + v.c0 + c : b * v.bs + static_cast<long long>(v.c0 + c) * S + s;
}

struct f4 {
    float v[4];
};
template <int V>
__device__ __forceinline__ f4 ldv(const float* p, long long i) {
    f4 r;
    if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p + i);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = p[i];
    }
    return r;
}
template <int V>
__device__ __forceinline__ f4 ldv(const __nv_bfloat16* p, long long i) {
    f4 r;
    if (V == 4) {
        const uint2 t = *reinterpret_cast<const uint2*>(p + i);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    } else {
        r.v[0] = __bfloat162float(p[i]);
    }
    return r;
}
template <int V>
__device__ __forceinline__ void stv(float* p, long long i, const f4& r) {
    if (V == 4) *reinterpret_cast<float4*>(p + i) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else p[i] = r.v[0];
}
template <int V>
__device__ __forceinline__ void stv(__nv_bfloat16* p, long long i, const f4& r) {
    if (V == 4) {
        uint2 t;
        *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(r.v[0], r.v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(r.v[2], r.v[3]);
        *reinterpret_cast<uint2*>(p + i) = t;
    } else {
        p[i] = __float2bfloat16_rn(r.v[0]);
    }
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// one thread per V consecutive elements along the memory-fastest index (s for channel-major, c for channels-last);
// 32-bit index math (the host checks B * 2C * S < 2^31)
template <int V>
__device__ __forceinline__ bool decode(int cl, int B, int C, int S, int& b, int& c, int& s) {
    const unsigned e = (blockIdx.x * kThreads + threadIdx.x) * V;
    const unsigned per_b = static_cast<unsigned>(C) * S;
    if (e >= per_b * static_cast<unsigned>(B)) return false;
    b = e / per_b;
    const unsigned r = e - b * per_b;
    if (cl) {
        s = r / C;
        c = r - s * C;
    } else {
        c = r / S;
        s = r - c * S;
    }
    return true;
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
gate_fwd_kernel(const T* __restrict__ g, const float* __restrict__ h, const float* __restrict__ x, float* __restrict__ xhr,
                View vg, View vh, View vx, View vo, int cl, int B, int C, int S) {
    int b, c, s;
    if (!decode<V>(cl, B, C, S, b, c, s)) return;
    View vg1 = vg, vo1 = vo;
    vg1.c0 = C;
    vo1.c0 = C;
    const f4 gr = ldv<V>(g, at(vg1, cl, S, b, c, s)), hv = ldv<V>(h, at(vh, cl, S, b, c, s));
    f4 hr;
#pragma unroll
    for (int i = 0; i < V; ++i) hr.v[i] = __fmul_rn(hv.v[i], sigmoidf(gr.v[i]));
    stv<V>(xhr, at(vo, cl, S, b, c, s), ldv<V>(x, at(vx, cl, S, b, c, s)));
    stv<V>(xhr, at(vo1, cl, S, b, c, s), hr);
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
gate_bwd_kernel(const float* __restrict__ dxhr, const T* __restrict__ g, const float* __restrict__ h, T* __restrict__ dg,
                float* __restrict__ dh, float* __restrict__ dx, View vd, View vg, View vh, View vo, int cl, int B, int C, int S) {
    int b, c, s;
    if (!decode<V>(cl, B, C, S, b, c, s)) return;
    View vg1 = vg, vd1 = vd;
    vg1.c0 = C;
    vd1.c0 = C;
    const f4 gr = ldv<V>(g, at(vg1, cl, S, b, c, s)), hv = ldv<V>(h, at(vh, cl, S, b, c, s));
    const f4 dhr = ldv<V>(dxhr, at(vd1, cl, S, b, c, s));
    f4 o_dh, o_dg, zero;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float r = sigmoidf(gr.v[i]);
        o_dh.v[i] = dhr.v[i] * r;
        o_dg.v[i] = dhr.v[i] * hv.v[i] * r * (1.f - r);
        zero.v[i] = 0.f;
    }
    const long long o = at(vo, cl, S, b, c, s);
    stv<V>(dx, o, ldv<V>(dxhr, at(vd, cl, S, b, c, s)));
    stv<V>(dh, o, o_dh);
    stv<V>(dg, at(vg, cl, S, b, c, s), zero);                                  // the update half belongs to the out stage
    stv<V>(dg, at(vg1, cl, S, b, c, s), o_dg);
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
out_fwd_kernel(const T* __restrict__ o, const T* __restrict__ g, const float* __restrict__ h, float* __restrict__ hn, View vo,
               View vg, View vh, View vn, int cl, int B, int C, int S) {
    int b, c, s;
    if (!decode<V>(cl, B, C, S, b, c, s)) return;
    const f4 gu = ldv<V>(g, at(vg, cl, S, b, c, s)), ov = ldv<V>(o, at(vo, cl, S, b, c, s)), hv = ldv<V>(h, at(vh, cl, S, b, c, s));
    f4 r;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float u = sigmoidf(gu.v[i]), cc = tanhf(ov.v[i]);
        r.v[i] = __fadd_rn(__fmul_rn(hv.v[i], 1.f - u), __fmul_rn(cc, u));     // h (1 - u) + c u, rounded like torch
    }
    stv<V>(hn, at(vn, cl, S, b, c, s), r);
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads)
out_bwd_kernel(const float* __restrict__ ghn, const T* __restrict__ o, const T* __restrict__ g, const float* __restrict__ h,
               T* __restrict__ d_o, T* __restrict__ dg, float* __restrict__ dh, View vn, View vo, View vg, View vh, View vd,
               int cl, int B, int C, int S) {
    int b, c, s;
    if (!decode<V>(cl, B, C, S, b, c, s)) return;
    const f4 gu = ldv<V>(g, at(vg, cl, S, b, c, s)), ov = ldv<V>(o, at(vo, cl, S, b, c, s)), hv = ldv<V>(h, at(vh, cl, S, b, c, s));
    const f4 gh = ldv<V>(ghn, at(vn, cl, S, b, c, s));
    f4 o_dh, o_do, o_dg, zero;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float u = sigmoidf(gu.v[i]), cc = tanhf(ov.v[i]);
        o_dh.v[i] = gh.v[i] * (1.f - u);
        o_do.v[i] = gh.v[i] * u * (1.f - cc * cc);
        o_dg.v[i] = gh.v[i] * (cc - hv.v[i]) * u * (1.f - u);
        zero.v[i] = 0.f;
    }
    View vg1 = vg;
    vg1.c0 = C;
    stv<V>(dh, at(vd, cl, S, b, c, s), o_dh);
    stv<V>(d_o, at(vo, cl, S, b, c, s), o_do);
    stv<V>(dg, at(vg, cl, S, b, c, s), o_dg);
    stv<V>(dg, at(vg1, cl, S, b, c, s), zero);                                 // the reset half belongs to the gate stage
}

inline View dense(int CC, int S) { return View{static_cast<long long>(CC) * S, CC, 0}; }

// vector width: 4 when the memory-fastest extent, the batch strides and the pointers allow 16-byte (fp32) accesses
inline int vec_width(int cl, int C, int S, long long h_bs, long long x_bs, std::initializer_list<const void*> ptrs) {
    const int inner = cl ? C : S;
    if (inner % 4 || h_bs % 4 || x_bs % 4) return 1;
    for (const void* p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) & 15u) return 1;
    return 4;
}

}  // namespace gru
}  // namespace forge

#define FORGE_GRU_PROLOGUE(fn)                                                                                            \
    if (B <= 0 || C <= 0 || S <= 0) return fail(fn, "non-positive size");                                                 \
    if (static_cast<long long>(B) * 2 * C * S >= 0x7fffffffLL) return fail(fn, "tensor too large (B * 2C * S must stay below 2^31)"); \
    cudaStream_t st_ = static_cast<cudaStream_t>(stream)
#define FORGE_GRU_BLOCKS(V) static_cast<unsigned>((static_cast<long long>(B) * C * S / (V) + kThreads - 1) / kThreads)
// dispatch on (gate dtype, vector width)
#define FORGE_GRU_DISPATCH(KERNEL, BF16, V, ...)                                                                          \
    do {                                                                                                                  \
        if (BF16) {                                                                                                       \
            typedef __nv_bfloat16 GT;                                                                                     \
            if ((V) == 4) KERNEL<GT, 4><<<FORGE_GRU_BLOCKS(4), kThreads, 0, st_>>>(__VA_ARGS__);                          \
            else KERNEL<GT, 1><<<FORGE_GRU_BLOCKS(1), kThreads, 0, st_>>>(__VA_ARGS__);                                   \
        } else {                                                                                                          \
            typedef float GT;                                                                                             \
            if ((V) == 4) KERNEL<GT, 4><<<FORGE_GRU_BLOCKS(4), kThreads, 0, st_>>>(__VA_ARGS__);                          \
            else KERNEL<GT, 1><<<FORGE_GRU_BLOCKS(1), kThreads, 0, st_>>>(__VA_ARGS__);                                   \
        }                                                                                                                 \
    } while (0)

extern "C" int forge_gru_gate_fwd(const void* g, int g_bf16, const float* h, long long h_bs, const float* x, long long x_bs,
                                  float* xhr, int channels_last, int B, int C, int S, void* stream) {
    FORGE_RANGE("forge_gru_gate_fwd");
    using namespace forge;
    using namespace forge::gru;
    const char* fn = "forge_gru_gate_fwd";
    if (!g || !h || !x || !xhr) return fail(fn, "null pointer");
    FORGE_GRU_PROLOGUE(fn);
    const View vg = dense(2 * C, S), vh{h_bs, C, 0}, vx{x_bs, C, 0}, vo = dense(2 * C, S);
    const int V = vec_width(channels_last, C, S, h_bs, x_bs, {g, h, x, xhr});
    FORGE_GRU_DISPATCH(gate_fwd_kernel, g_bf16, V, static_cast<const GT*>(g), h, x, xhr, vg, vh, vx, vo, channels_last, B, C, S);
    return check_launch(fn);
}

extern "C" int forge_gru_gate_bwd(const float* d_xhr, const void* g, int g_bf16, const float* h, long long h_bs, void* d_g,
                                  float* d_h, float* d_x, int channels_last, int B, int C, int S, void* stream) {
    FORGE_RANGE("forge_gru_gate_bwd");
    using namespace forge;
    using namespace forge::gru;
    const char* fn = "forge_gru_gate_bwd";
    if (!d_xhr || !g || !h || !d_g || !d_h || !d_x) return fail(fn, "null pointer");
    FORGE_GRU_PROLOGUE(fn);
    const View vd = dense(2 * C, S), vg = dense(2 * C, S), vh{h_bs, C, 0}, vo = dense(C, S);
    const int V = vec_width(channels_last, C, S, h_bs, 0, {d_xhr, g, h, d_g, d_h, d_x});
    FORGE_GRU_DISPATCH(gate_bwd_kernel, g_bf16, V, d_xhr, static_cast<const GT*>(g), h, static_cast<GT*>(d_g), d_h, d_x, vd, vg, vh,
                       vo, channels_last, B, C, S);
    return check_launch(fn);
}

extern "C" int forge_gru_out_fwd(const void* o, const void* g, int og_bf16, const float* h, long long h_bs, float* h_new,
                                 int channels_last, int B, int C, int S, void* stream) {
    FORGE_RANGE("forge_gru_out_fwd");
    using namespace forge;
    using namespace forge::gru;
    const char* fn = "forge_gru_out_fwd";
    if (!o || !g || !h || !h_new) return fail(fn, "null pointer");
    FORGE_GRU_PROLOGUE(fn);
    const View vo = dense(C, S), vg = dense(2 * C, S), vh{h_bs, C, 0}, vn = dense(C, S);
    const int V = vec_width(channels_last, C, S, h_bs, 0, {o, g, h, h_new});
    FORGE_GRU_DISPATCH(out_fwd_kernel, og_bf16, V, static_cast<const GT*>(o), static_cast<const GT*>(g), h, h_new, vo, vg, vh, vn,
                       channels_last, B, C, S);
    return check_launch(fn);
}

extern "C" int forge_gru_out_bwd(const float* d_h_new, const void* o, const void* g, int og_bf16, const float* h, long long h_bs,
                                 void* d_o, void* d_g, float* d_h, int channels_last, int B, int C, int S, void* stream) {
    FORGE_RANGE("forge_gru_out_bwd");
    using namespace forge;
    using namespace forge::gru;
    const char* fn = "forge_gru_out_bwd";
    if (!d_h_new || !o || !g || !h || !d_o || !d_g || !d_h) return fail(fn, "null pointer");
    FORGE_GRU_PROLOGUE(fn);
    const View vn = dense(C, S), vo = dense(C, S), vg = dense(2 * C, S), vh{h_bs, C, 0}, vd = dense(C, S);
    const int V = vec_width(channels_last, C, S, h_bs, 0, {d_h_new, o, g, h, d_o, d_g, d_h});
    FORGE_GRU_DISPATCH(out_bwd_kernel, og_bf16, V, d_h_new, static_cast<const GT*>(o), static_cast<const GT*>(g), h,
                       static_cast<GT*>(d_o), static_cast<GT*>(d_g), d_h, vn, vo, vg, vh, vd, channels_last, B, C, S);
    return check_launch(fn);
}
