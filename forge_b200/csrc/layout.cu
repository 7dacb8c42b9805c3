// libforge_b200: error plumbing, [n][C][S] <-> [n][S][C] re-layout kernels, sampler test hook.
#include "common.cuh"

namespace forge {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(const char* fn, const std::string& msg) {
    g_last_error = std::string(fn) + ": " + msg;
    return 1;
}

int check_launch(const char* fn) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(fn, std::string("CUDA launch failed: ") + cudaGetErrorString(e));
    return 0;
}

// ---- transposes ---------------------------------------------------------------------------------
// One CTA moves a [C_TILE x 32] tile through shared memory so that both the read (along S) and the
// write (along C) are coalesced.  HBM-bound: algorithmic bytes = 2 * n * C * S * 4.
constexpr int kTile = 32;

__global__ void __launch_bounds__(256) ncs_to_nsc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                                         long long S) {
    __shared__ float tile[kTile][kTile + 1];
    const long long s0 = static_cast<long long>(blockIdx.x) * kTile;
    const int c0 = blockIdx.y * kTile;
    const float* sp = src + static_cast<long long>(blockIdx.z) * C * S;
    float* dp = dst + static_cast<long long>(blockIdx.z) * C * S;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const int c = c0 + r;
        const long long s = s0 + tx;
        if (c < C && s < S) tile[r][tx] = sp[static_cast<long long>(c) * S + s];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const long long s = s0 + r;
        const int c = c0 + tx;
        if (c < C && s < S) dp[s * C + c] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256) nsc_to_ncs_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                                         long long S) {
    __shared__ float tile[kTile][kTile + 1];
    const long long s0 = static_cast<long long>(blockIdx.x) * kTile;
    const int c0 = blockIdx.y * kTile;
    const float* sp = src + static_cast<long long>(blockIdx.z) * C * S;
    float* dp = dst + static_cast<long long>(blockIdx.z) * C * S;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const long long s = s0 + r;
        const int c = c0 + tx;
        if (c < C && s < S) tile[r][tx] = sp[s * C + c];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const int c = c0 + r;
        const long long s = s0 + tx;
        if (c < C && s < S) dp[static_cast<long long>(c) * S + s] = tile[tx][r];
    }
}

__global__ void sample_points_kernel(const float* __restrict__ pts, int M, int D, int H, int W, int ac,
                                     int* __restrict__ base, unsigned char* __restrict__ mask) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float x = pts[3 * m], y = pts[3 * m + 1], z = pts[3 * m + 2];
    const float ix = ac ? unnormalize_ac(x, W) : unnormalize_nac(x, W);
    const float iy = ac ? unnormalize_ac(y, H) : unnormalize_nac(y, H);
    const float iz = ac ? unnormalize_ac(z, D) : unnormalize_nac(z, D);
    const Tri t = make_tri(ix, iy, iz, D, H, W);
    base[3 * m] = t.x0;
    base[3 * m + 1] = t.y0;
    base[3 * m + 2] = t.z0;
    mask[m] = static_cast<unsigned char>(t.mask);
}

static int transpose_common(const char* fn, bool to_nsc, const float* src, float* dst, int n, int C, long long S,
                            void* stream) {
    if (!src || !dst) return fail(fn, "null pointer");
    if (n <= 0 || C <= 0 || S <= 0) return fail(fn, "non-positive size");
    if (n > 65535) return fail(fn, "n exceeds grid.z limit");
    const long long sx = (S + kTile - 1) / kTile;
    const int cy = (C + kTile - 1) / kTile;
    if (sx > 2147483647LL || cy > 65535) return fail(fn, "tensor too large for one launch");
    dim3 grid(static_cast<unsigned>(sx), cy, n);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (to_nsc)
        ncs_to_nsc_kernel<<<grid, 256, 0, st>>>(src, dst, C, S);
    else
        nsc_to_ncs_kernel<<<grid, 256, 0, st>>>(src, dst, C, S);
    return check_launch(fn);
}

}  // namespace forge

extern "C" {

int forge_abi_version(void) { return FORGE_ABI_VERSION; }

const char* forge_last_error(void) { return forge::g_last_error.c_str(); }

int forge_ncs_to_nsc(const float* src, float* dst, int n, int C, long long S, void* stream) {
    return forge::transpose_common("forge_ncs_to_nsc", true, src, dst, n, C, S, stream);
}

int forge_nsc_to_ncs(const float* src, float* dst, int n, int C, long long S, void* stream) {
    return forge::transpose_common("forge_nsc_to_ncs", false, src, dst, n, C, S, stream);
}

int forge_sample_points(const float* pts, int M, int D, int H, int W, int align_corners, int* base,
                        unsigned char* mask, void* stream) {
    using namespace forge;
    if (!pts || !base || !mask) return fail("forge_sample_points", "null pointer");
    if (M <= 0 || D <= 0 || H <= 0 || W <= 0) return fail("forge_sample_points", "non-positive size");
    sample_points_kernel<<<(M + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, M, D, H, W,
                                                                                         align_corners, base, mask);
    return check_launch("forge_sample_points");
}

}  // extern "C"
