// K2: affine resample of feature volumes into the canonical camera frame (forward + backward).
//
// Replaces reference models/rotate.py:127-141: the voxel-centre grid is affine in (w, h, d), so it
// is evaluated in registers and never materialised; volumes are channels-last so the 8 corner
// fetches of an output voxel are contiguous C*4-byte runs (512 B for C = 128) and the output write
// is one contiguous run -- every HBM byte moved is an algorithmic byte.  HBM-bound.
//
// Work mapping: one thread per (output voxel, channel vector).  A CTA walks a compact
// kTz x kTy x kTx block of output voxels so that the rotated input footprint stays L1/L2-hot.
#include "common.cuh"

namespace forge {

constexpr int kRotThreads = 256;
constexpr int kTx = 8, kTy = 4, kTz = 4;            // output-voxel block per CTA
constexpr int kTileVox = kTx * kTy * kTz;           // 128

struct RotJob {
    int src, dst, kind;
};

// sample position (unnormalised, align_corners=False) of output voxel (d, h, w) under job affine
__device__ __forceinline__ Tri rotate_tri(const float* __restrict__ A, float gxw, float gyh, float gzd, float inv_max,
                                          int D, int H, int W) {
    // [g, 1] @ T^T, accumulated in k order like a GEMM inner product
    const float cx = fmaf(1.f, A[3], fmaf(gzd, A[2], fmaf(gyh, A[1], __fmul_rn(gxw, A[0]))));
    const float cy = fmaf(1.f, A[7], fmaf(gzd, A[6], fmaf(gyh, A[5], __fmul_rn(gxw, A[4]))));
    const float cz = fmaf(1.f, A[11], fmaf(gzd, A[10], fmaf(gyh, A[9], __fmul_rn(gxw, A[8]))));
    // tensor / python-scalar on CUDA multiplies by the reciprocal (ATen div_true_kernel_cuda)
    const float sx = __fmul_rn(cx, inv_max), sy = __fmul_rn(cy, inv_max), sz = __fmul_rn(cz, inv_max);
    return make_tri(unnormalize_nac(sx, W), unnormalize_nac(sy, H), unnormalize_nac(sz, D), D, H, W);
}

template <typename VecT>
__device__ __forceinline__ VecT vzero();
template <>
__device__ __forceinline__ float vzero<float>() {
    return 0.f;
}
template <>
__device__ __forceinline__ float4 vzero<float4>() {
    return make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void vfma(float& a, float v, float w) { a = fmaf(v, w, a); }
__device__ __forceinline__ void vfma(float4& a, const float4& v, float w) {
    a.x = fmaf(v.x, w, a.x);
    a.y = fmaf(v.y, w, a.y);
    a.z = fmaf(v.z, w, a.z);
    a.w = fmaf(v.w, w, a.w);
}

// CU = channel vectors per voxel (C/4 for float4, C for float)
template <typename VecT>
__global__ void __launch_bounds__(kRotThreads)
rotate_fwd_kernel(const VecT* __restrict__ in, const float* __restrict__ affine, const int* __restrict__ jobs,
                  const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                  float inv_max, VecT* __restrict__ out, int CU, int D, int H, int W, int tiles_x, int tiles_y) {
    __shared__ float A[12];
    const int m = blockIdx.y;
    const RotJob job = {jobs[3 * m], jobs[3 * m + 1], jobs[3 * m + 2]};
    if (threadIdx.x < 12) A[threadIdx.x] = affine[12 * m + threadIdx.x];
    __syncthreads();
    const int tz = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - tz * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const long long vol = static_cast<long long>(D) * H * W;
    const VecT* src = in + static_cast<long long>(job.src) * vol * CU;
    VecT* dst = out + static_cast<long long>(job.dst) * vol * CU;

    for (int e = threadIdx.x; e < kTileVox * CU; e += kRotThreads) {
        const int v = e / CU, cu = e - v * CU;
        const int w = tx * kTx + (v % kTx), h = ty * kTy + ((v / kTx) % kTy), d = tz * kTz + v / (kTx * kTy);
        if (w >= W || h >= H || d >= D) continue;
        const long long o = (static_cast<long long>(d) * H + h) * W + w;
        if (job.kind == 1) {   // view-0 passthrough (models/rotate.py:141)
            dst[o * CU + cu] = src[o * CU + cu];
            continue;
        }
        const Tri t = rotate_tri(A, gx[w], gy[h], gz[d], inv_max, D, H, W);
        VecT acc = vzero<VecT>();
#pragma unroll
        for (int cn = 0; cn < 8; ++cn) {
            if ((t.mask >> cn) & 1u) {
                const long long vox =
                    (static_cast<long long>(t.z0 + (cn >> 2)) * H + (t.y0 + ((cn >> 1) & 1))) * W + (t.x0 + (cn & 1));
                vfma(acc, __ldg(src + vox * CU + cu), tri_weight(t, cn));
            }
        }
        dst[o * CU + cu] = acc;
    }
}

static int rotate_check(const char* fn, const void* vox, const void* aff, const void* jobs, const void* gx,
                        const void* gy, const void* gz, float gmax, int M, int C, int D, int H, int W) {
    if (!vox || !aff || !jobs || !gx || !gy || !gz) return fail(fn, "null pointer");
    if (M <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(fn, "non-positive size");
    if (!(gmax > 0.f)) return fail(fn, "grid_coord_max must be positive");
    if (M > 65535) return fail(fn, "more than 65535 jobs in one launch");
    return 0;
}

}  // namespace forge

extern "C" int forge_rotate_fwd(const float* vox_cl, const float* affine12, const int* jobs, const float* gx,
                                const float* gy, const float* gz, float grid_coord_max, float* out_cl, int M, int C,
                                int D, int H, int W, void* stream) {
    using namespace forge;
    const char* fn = "forge_rotate_fwd";
    if (int e = rotate_check(fn, vox_cl, affine12, jobs, gx, gy, gz, grid_coord_max, M, C, D, H, W)) return e;
    if (!out_cl) return fail(fn, "null pointer");
    const int tiles_x = (W + kTx - 1) / kTx, tiles_y = (H + kTy - 1) / kTy, tiles_z = (D + kTz - 1) / kTz;
    dim3 grid(tiles_x * tiles_y * tiles_z, M);
    const float inv_max = 1.0f / grid_coord_max;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (C % 4 == 0 && aligned16(vox_cl) && aligned16(out_cl)) {
        rotate_fwd_kernel<float4><<<grid, kRotThreads, 0, st>>>(reinterpret_cast<const float4*>(vox_cl), affine12, jobs,
                                                               gx, gy, gz, inv_max, reinterpret_cast<float4*>(out_cl),
                                                               C / 4, D, H, W, tiles_x, tiles_y);
    } else {
        rotate_fwd_kernel<float><<<grid, kRotThreads, 0, st>>>(vox_cl, affine12, jobs, gx, gy, gz, inv_max, out_cl, C, D,
                                                              H, W, tiles_x, tiles_y);
    }
    return check_launch(fn);
}

extern "C" int forge_rotate_bwd(const float*, const float*, const int*, const float*, const float*, const float*, float,
                                const float*, float*, float*, int, int, int, int, int, void*) {
    return forge::fail("forge_rotate_bwd", "not implemented yet");
}
