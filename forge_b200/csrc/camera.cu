// Host-glue kernels: the per-view camera / pose algebra that feeds K1 and K2, one launch each instead of
// the 20-40 tiny torch kernels (and the batched LU inverse) the Python mirrors used to spend ~0.2-0.4 ms on.
//
//   forge_camera_prep_fwd/bwd   OpenCV (R, t, K_half) -> cam12 = (o_local, M_local) consumed by the raymarcher
//                               (= PyTorch3D cameras_from_opencv_projection + NDC un-projection +
//                               Volumes.world_to_local, reference models/volume_render.py:53-61) and the pixel
//                               position of the world origin (reference :77-83, :97-103)
//   forge_pose_affine_fwd       A[b, v] = (pose[b, 0] @ inverse(pose[b, v]))[:3, :] for v >= 1, identity for v = 0
//                               (reference models/rotate.py:64-89), plus inverse(pose) for the backward pass
//
// One thread per view; everything is a few dozen flops, so the only figure of merit is "one launch".
#include "common.cuh"

namespace forge {

struct Cam {
    float R[9], t[3], fx, fy, cx, cy;
};

__device__ __forceinline__ Cam load_cam(const float* R, const float* T, const float* K, int n) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = R[9 * n + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = T[3 * n + i];
    c.fx = K[9 * n + 0];
    c.cx = K[9 * n + 2];
    c.fy = K[9 * n + 4];
    c.cy = K[9 * n + 5];
    return c;
}

// sign-preserving clamp of t_z away from 0 (PyTorch3D transform_points eps): returns the divisor and whether
// the clamp is inactive (gradient passes)
__device__ __forceinline__ float clamp_tz(float tz, float eps, bool* pass) {
    const float s = tz > 0.f ? 1.f : (tz < 0.f ? -1.f : 1.f);
    *pass = fabsf(tz) > eps;
    return s * fmaxf(fabsf(tz), eps);
}

__global__ void camera_prep_fwd_kernel(const float* __restrict__ R, const float* __restrict__ T, const float* __restrict__ K,
                                       int N, float sx, float sy, float sz, float eps, float* __restrict__ cam12,
                                       float* __restrict__ oproj) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const Cam c = load_cam(R, T, K, n);
    const float s[3] = {sx, sy, sz};
    float* o = cam12 + 12 * n;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // column i of R = row i of R^T
        const float r0 = c.R[0 + i], r1 = c.R[3 + i], r2 = c.R[6 + i];
        o[i] = -(r0 * c.t[0] + r1 * c.t[1] + r2 * c.t[2]) / s[i];
        // M = diag(1/s) R^T K^-1,  K^-1 = [[1/fx, 0, -cx/fx], [0, 1/fy, -cy/fy], [0, 0, 1]]
        o[3 + 3 * i + 0] = (r0 / c.fx) / s[i];
        o[3 + 3 * i + 1] = (r1 / c.fy) / s[i];
        o[3 + 3 * i + 2] = (r0 * (-c.cx / c.fx) + r1 * (-c.cy / c.fy) + r2) / s[i];
    }
    if (oproj) {
        bool pass;
        const float tz = clamp_tz(c.t[2], eps, &pass);
        oproj[2 * n + 0] = c.fx * c.t[0] / tz + c.cx;
        oproj[2 * n + 1] = c.fy * c.t[1] / tz + c.cy;
    }
}

__global__ void camera_prep_bwd_kernel(const float* __restrict__ R, const float* __restrict__ T, const float* __restrict__ K,
                                       int N, float sx, float sy, float sz, float eps, const float* __restrict__ g_cam12,
                                       const float* __restrict__ g_oproj, float* __restrict__ gR, float* __restrict__ gT,
                                       float* __restrict__ gK) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const Cam c = load_cam(R, T, K, n);
    const float s[3] = {sx, sy, sz};
    float dR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dt[3] = {0.f, 0.f, 0.f};
    float dfx = 0.f, dfy = 0.f, dcx = 0.f, dcy = 0.f;
    if (g_cam12) {
        const float* g = g_cam12 + 12 * n;
        const float ifx = 1.f / c.fx, ify = 1.f / c.fy;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float is = 1.f / s[i];
            const float go = g[i] * is, g0 = g[3 + 3 * i] * is, g1 = g[3 + 3 * i + 1] * is, g2 = g[3 + 3 * i + 2] * is;
            const float r0 = c.R[0 + i], r1 = c.R[3 + i];
            // o_i = -(sum_j R[j][i] t_j) / s_i
            dR[0 + i] += -go * c.t[0];
            dR[3 + i] += -go * c.t[1];
            dR[6 + i] += -go * c.t[2];
            dt[0] += -go * r0;
            dt[1] += -go * r1;
            dt[2] += -go * c.R[6 + i];
            // M_i0 = r0 / fx, M_i1 = r1 / fy, M_i2 = -r0 cx / fx - r1 cy / fy + r2   (each / s_i)
            dR[0 + i] += g0 * ifx - g2 * c.cx * ifx;
            dR[3 + i] += g1 * ify - g2 * c.cy * ify;
            dR[6 + i] += g2;
            dfx += (-g0 * r0 + g2 * r0 * c.cx) * ifx * ifx;
            dfy += (-g1 * r1 + g2 * r1 * c.cy) * ify * ify;
            dcx += -g2 * r0 * ifx;
            dcy += -g2 * r1 * ify;
        }
    }
    if (g_oproj) {
        bool pass;
        const float tz = clamp_tz(c.t[2], eps, &pass);
        const float gu = g_oproj[2 * n], gv = g_oproj[2 * n + 1], itz = 1.f / tz;
        dt[0] += gu * c.fx * itz;
        dt[1] += gv * c.fy * itz;
        if (pass) dt[2] += -(gu * c.fx * c.t[0] + gv * c.fy * c.t[1]) * itz * itz;
        dfx += gu * c.t[0] * itz;
        dfy += gv * c.t[1] * itz;
        dcx += gu;
        dcy += gv;
    }
    if (gR) {
#pragma unroll
        for (int i = 0; i < 9; ++i) gR[9 * n + i] = dR[i];
    }
    if (gT) {
#pragma unroll
        for (int i = 0; i < 3; ++i) gT[3 * n + i] = dt[i];
    }
    if (gK) {
        float* k = gK + 9 * n;
        k[0] = dfx; k[1] = 0.f; k[2] = dcx;
        k[3] = 0.f; k[4] = dfy; k[5] = dcy;
        k[6] = 0.f; k[7] = 0.f; k[8] = 0.f;
    }
}

// general 4x4 inverse by cofactors, evaluated in fp64 and rounded once (the reference inverts in fp32 with LU;
// both agree with the exact inverse to fp32 rounding for the well-conditioned rigid poses of the path)
__device__ __forceinline__ bool inverse4(const double* m, double* inv) {
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0) return false;
    const double id = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 16; ++i) inv[i] *= id;
    return true;
}

__global__ void pose_affine_fwd_kernel(const float* __restrict__ poses, int B, int t, float* __restrict__ affine12,
                                       float* __restrict__ pose_inv, int* __restrict__ singular) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * t) return;
    const int b = e / t, v = e - b * t;
    double m[16], inv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = static_cast<double>(poses[16 * e + i]);
    const bool ok = inverse4(m, inv);
    if (!ok) {
        if (singular) atomicExch(singular, 1);
#pragma unroll
        for (int i = 0; i < 16; ++i) inv[i] = __longlong_as_double(0x7ff8000000000000LL);     // NaN, like a failed LU
    }
    if (pose_inv) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pose_inv[16 * e + i] = static_cast<float>(inv[i]);
    }
    float* A = affine12 + 12 * e;
    if (v == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) A[i] = (i == 0 || i == 5 || i == 10) ? 1.f : 0.f;
        return;
    }
    // fp32 product with the rounded inverse, like the reference's fp32 matmul
    const float* p0 = poses + 16 * (b * t);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fmaf(p0[4 * r + k], static_cast<float>(inv[4 * k + c]), acc);
            A[4 * r + c] = acc;
        }
}

}  // namespace forge

extern "C" int forge_camera_prep_fwd(const float* R, const float* T, const float* K_half, int N, float sx, float sy, float sz,
                                     float eps, float* cam12, float* origin_proj, void* stream) {
    FORGE_RANGE("forge_camera_prep_fwd");
    using namespace forge;
    const char* fn = "forge_camera_prep_fwd";
    if (!R || !T || !K_half || !cam12) return fail(fn, "null pointer");
    if (N <= 0) return fail(fn, "non-positive size");
    if (!(sx > 0.f) || !(sy > 0.f) || !(sz > 0.f)) return fail(fn, "volume scale must be positive (volume sides must exceed 1 voxel)");
    camera_prep_fwd_kernel<<<(N + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(R, T, K_half, N, sx, sy, sz, eps, cam12,
                                                                                        origin_proj);
    return check_launch(fn);
}

extern "C" int forge_camera_prep_bwd(const float* R, const float* T, const float* K_half, int N, float sx, float sy, float sz,
                                     float eps, const float* g_cam12, const float* g_origin_proj, float* grad_R, float* grad_T,
                                     float* grad_K, void* stream) {
    FORGE_RANGE("forge_camera_prep_bwd");
    using namespace forge;
    const char* fn = "forge_camera_prep_bwd";
    if (!R || !T || !K_half) return fail(fn, "null pointer");
    if (N <= 0) return fail(fn, "non-positive size");
    camera_prep_bwd_kernel<<<(N + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(R, T, K_half, N, sx, sy, sz, eps, g_cam12,
                                                                                        g_origin_proj, grad_R, grad_T, grad_K);
    return check_launch(fn);
}

extern "C" int forge_pose_affine_fwd(const float* poses, int B, int t, float* affine12, float* pose_inv, int* singular_flag,
                                     void* stream) {
    FORGE_RANGE("forge_pose_affine_fwd");
    using namespace forge;
    const char* fn = "forge_pose_affine_fwd";
    if (!poses || !affine12) return fail(fn, "null pointer");
    if (B <= 0 || t <= 0) return fail(fn, "non-positive size");
    pose_affine_fwd_kernel<<<(B * t + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(poses, B, t, affine12, pose_inv,
                                                                                            singular_flag);
    return check_launch(fn);
}

// ---- x2 bilinear upsample of the silhouette / depth maps ---------------------------------------------
// F.upsample(mode='bilinear') = interpolate(align_corners=False) of reference models/volume_render.py:69,74 for
// exactly twice the size: source coordinate max(0.5 (o + 0.5) - 0.5, 0), i.e. weights 0.25 / 0.75 with the
// borders clamped.  One launch handles up to two maps (silhouette and depth); a thread makes a 2x2 output quad.
namespace forge {

__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float* __restrict__ src0, const float* __restrict__ src1, float* __restrict__ dst0,
                      float* __restrict__ dst1, int M, int Sh, int Sw) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_map = static_cast<long long>(M) * Sh * Sw;
    if (e >= per_map * (src1 ? 2 : 1)) return;
    const bool second = e >= per_map;
    const long long r = second ? e - per_map : e;
    const float* src = second ? src1 : src0;
    float* dst = second ? dst1 : dst0;
    const int x = static_cast<int>(r % Sw), y = static_cast<int>((r / Sw) % Sh);
    const long long m = r / (static_cast<long long>(Sw) * Sh);
    const float* s = src + m * Sh * Sw;
    const int xm = max(x - 1, 0), xp = min(x + 1, Sw - 1), ym = max(y - 1, 0), yp = min(y + 1, Sh - 1);
    const float a = s[ym * Sw + xm], b = s[ym * Sw + x], c = s[ym * Sw + xp];
    const float d = s[y * Sw + xm], f = s[y * Sw + x], g = s[y * Sw + xp];
    const float h = s[yp * Sw + xm], i = s[yp * Sw + x], j = s[yp * Sw + xp];
    // output (2y, 2x): rows (y-1: 0.25, y: 0.75), cols (x-1: 0.25, x: 0.75); at y = 0 / x = 0 the clamp makes both taps equal
    const float wl0 = x > 0 ? 0.25f : 0.f, wl1 = 1.f - wl0;            // left output column: weights of (x-1, x)
    const float wr1 = 0.25f, wr0 = 0.75f;                             // right output column: weights of (x, x+1)
    const float wt0 = y > 0 ? 0.25f : 0.f, wt1 = 1.f - wt0;
    const float top_l = wl0 * a + wl1 * b, top_r = wr0 * b + wr1 * c;
    const float mid_l = wl0 * d + wl1 * f, mid_r = wr0 * f + wr1 * g;
    const float bot_l = wl0 * h + wl1 * i, bot_r = wr0 * i + wr1 * j;
    float* o = dst + (m * 2 * Sh + 2 * y) * (2 * Sw) + 2 * x;
    *reinterpret_cast<float2*>(o) = make_float2(wt0 * top_l + wt1 * mid_l, wt0 * top_r + wt1 * mid_r);
    *reinterpret_cast<float2*>(o + 2 * Sw) = make_float2(0.75f * mid_l + 0.25f * bot_l, 0.75f * mid_r + 0.25f * bot_r);
}

// adjoint: every source pixel gathers its (up to) 4x4 output neighbourhood with the same weights
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float* __restrict__ g0, const float* __restrict__ g1, float* __restrict__ gs0,
                      float* __restrict__ gs1, int M, int Sh, int Sw) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_map = static_cast<long long>(M) * Sh * Sw;
    if (e >= per_map * (g1 ? 2 : 1)) return;
    const bool second = e >= per_map;
    const long long r = second ? e - per_map : e;
    const float* g = second ? g1 : g0;
    float* gs = second ? gs1 : gs0;
    const int x = static_cast<int>(r % Sw), y = static_cast<int>((r / Sw) % Sh);
    const long long m = r / (static_cast<long long>(Sw) * Sh);
    const float* gm = g + m * 4 * Sh * Sw;
    const int OW = 2 * Sw, OH = 2 * Sh;
    // 1-D weights of source index x on output columns 2x-1 .. 2x+2 (clamped taps fold onto the border pixel)
    float wx[4], wy[4];
    wx[0] = 0.25f; wx[1] = 0.75f; wx[2] = 0.75f; wx[3] = 0.25f;
    wy[0] = 0.25f; wy[1] = 0.75f; wy[2] = 0.75f; wy[3] = 0.25f;
    if (x == 0) { wx[0] = 0.f; wx[1] = 1.f; }
    if (x == Sw - 1) { wx[3] = 0.f; wx[2] = 1.f; }
    if (y == 0) { wy[0] = 0.f; wy[1] = 1.f; }
    if (y == Sh - 1) { wy[3] = 0.f; wy[2] = 1.f; }
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
        const int oy = 2 * y - 1 + dy;
        if (oy < 0 || oy >= OH) continue;
        float row = 0.f;
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
            const int ox = 2 * x - 1 + dx;
            if (ox >= 0 && ox < OW) row = fmaf(wx[dx], gm[static_cast<long long>(oy) * OW + ox], row);
        }
        acc = fmaf(wy[dy], row, acc);
    }
    gs[r] = acc;
}

}  // namespace forge

extern "C" int forge_upsample2x_fwd(const float* src0, const float* src1, float* dst0, float* dst1, int M, int S_h, int S_w,
                                    void* stream) {
    FORGE_RANGE("forge_upsample2x_fwd");
    using namespace forge;
    const char* fn = "forge_upsample2x_fwd";
    if (!src0 || !dst0 || (src1 && !dst1)) return fail(fn, "null pointer");
    if (M <= 0 || S_h <= 0 || S_w <= 0) return fail(fn, "non-positive size");
    if ((reinterpret_cast<uintptr_t>(dst0) & 7u) || (dst1 && (reinterpret_cast<uintptr_t>(dst1) & 7u)))
        return fail(fn, "outputs must be 8-byte aligned");
    const long long total = static_cast<long long>(M) * S_h * S_w * (src1 ? 2 : 1);
    upsample2x_fwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src0, src1, dst0, dst1, M, S_h, S_w);
    return check_launch(fn);
}

extern "C" int forge_upsample2x_bwd(const float* g_dst0, const float* g_dst1, float* g_src0, float* g_src1, int M, int S_h,
                                    int S_w, void* stream) {
    FORGE_RANGE("forge_upsample2x_bwd");
    using namespace forge;
    const char* fn = "forge_upsample2x_bwd";
    if (!g_dst0 || !g_src0 || (g_dst1 && !g_src1)) return fail(fn, "null pointer");
    if (M <= 0 || S_h <= 0 || S_w <= 0) return fail(fn, "non-positive size");
    const long long total = static_cast<long long>(M) * S_h * S_w * (g_dst1 ? 2 : 1);
    upsample2x_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g_dst0, g_dst1, g_src0, g_src1, M, S_h, S_w);
    return check_launch(fn);
}
