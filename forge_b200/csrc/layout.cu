// libforge_b200: error plumbing, [n][C][S] <-> [n][S][C] re-layout kernels, sampler test hook.
#include <mutex>
#include <vector>

#include "tensormap.cuh"

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

namespace {
struct SmemOptIn {
    const void* kernel;
    int device;
    size_t bytes;
};
std::mutex g_attr_mutex;
std::vector<SmemOptIn> g_smem_optin;
std::vector<int> g_sm_count;     // indexed by device
}  // namespace

int ensure_dynamic_smem(const char* fn, const void* kernel, size_t bytes) {
    if (bytes + 2048 <= 48 * 1024) return 0;      // static shared memory of the kernel counts against the 48 KB default too
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(fn, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    for (SmemOptIn& o : g_smem_optin) {
        if (o.kernel == kernel && o.device == dev) {
            if (o.bytes >= bytes) return 0;
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
            if (e != cudaSuccess) return fail(fn, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
            o.bytes = bytes;
            return 0;
        }
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return fail(fn, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    g_smem_optin.push_back({kernel, dev, bytes});
    return 0;
}

int current_sm_count(const char* fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        fail(fn, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
        return 0;
    }
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    if (dev >= static_cast<int>(g_sm_count.size())) g_sm_count.resize(dev + 1, 0);
    if (g_sm_count[dev] == 0) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n <= 0) {
            fail(fn, std::string("cudaDeviceGetAttribute: ") + cudaGetErrorString(e));
            return 0;
        }
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}

int encode_tensor_map(const char* fn, CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                      const unsigned long long* dims, const unsigned long long* strides_bytes, const unsigned* box,
                      CUtensorMapSwizzle swizzle) {
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = [] {       // resolved once through the runtime (no link against libcuda)
        void* f = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<Encode>(f);
    }();
    if (!encode) return fail(fn, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gs[i - 1] = strides_bytes[i - 1];
    }
    const CUresult r = encode(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(fn, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
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

// ---- render-volume packing -------------------------------------------------------------------------
// forge_pack_volume: [V][16][D][H][W] (or channels-last [V][D][H][W][16]) features + [V][D][H][W] density -> zero-bordered
// channels-last feat_pad [V][D+2][H+2][W+2][16] and density quads [V][D+2][H+1][W+1][4].  HBM-bound: reads V*17*D*H*W*4
// bytes, writes ~1.1x that + 4x the density.
//
// Round 1 / first half of round 2 went through a shared-memory transpose (one CTA per group of padded y-rows, 128-bit
// loads, [row][x][channel] tile, load phase -> barrier -> store phase): 56.6 % of the HBM peak at cfg-2, 76 % at cfg-4
// with 4 rows per CTA.  The kernel below needs no shared memory: one thread per padded voxel gathers its 16 channels
// with 16 independent scalar loads (a warp reads 128 contiguous bytes per channel) and writes its 64-byte record with
// four 16-byte stores; one thread per density quad.  No barrier, no phases, fine-grained CTAs (short tail): 67.6 % at
// cfg-2 (0.044 -> 0.037 ms), 88.6 % at cfg-4 (0.511 -> 0.439 ms) -- profiles/r02_ab_pack_direct.jsonl.
constexpr int kPackThreads = 256;

// gradient of the padded feature volume back to the caller's layout (interior voxels only)
__global__ void __launch_bounds__(kPackThreads)
unpack_grad_kernel(const float* __restrict__ grad_pad, float* __restrict__ grad, int out_cl, int D, int H, int W) {
    extern __shared__ float sm[];
    float* tile = sm;   // [16][W + 1]
    const int Wp = W + 2, Hp = H + 2;
    const int v = blockIdx.y;
    const int z = blockIdx.x / H, y = blockIdx.x - z * H;
    const long long S = static_cast<long long>(D) * H * W;
    const float* src = grad_pad + (((static_cast<long long>(v) * (D + 2) + z + 1) * Hp + y + 1) * Wp + 1) * 16;
    if (out_cl) {
        float* dst = grad + (static_cast<long long>(v) * S + (static_cast<long long>(z) * H + y) * W) * 16;
        for (int e = threadIdx.x; e < W * 16; e += kPackThreads) dst[e] = src[e];
        return;
    }
    for (int e = threadIdx.x; e < W * 16; e += kPackThreads) tile[(e & 15) * (W + 1) + (e >> 4)] = src[e];
    __syncthreads();
    float* dst = grad + static_cast<long long>(v) * 16 * S + (static_cast<long long>(z) * H + y) * W;
    for (int e = threadIdx.x; e < 16 * W; e += kPackThreads) {
        const int ch = e / W, x = e - ch * W;
        dst[static_cast<long long>(ch) * S + x] = tile[ch * (W + 1) + x];
    }
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

namespace forge {

__global__ void __launch_bounds__(kPackThreads)
pack_volume_kernel(const float* __restrict__ feat, int feat_cl, const float* __restrict__ dens, float4* __restrict__ feat_pad,
                   float4* __restrict__ dens_quad, int D, int H, int W, long long n_vox, long long n_quad) {
    const int Wp = W + 2, Hp = H + 2, Wq = W + 1, Hq = H + 1;
    const long long S = static_cast<long long>(D) * H * W;
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i < n_vox) {
        const int xp = static_cast<int>(i % Wp), yp = static_cast<int>((i / Wp) % Hp);
        const long long vz = i / (static_cast<long long>(Wp) * Hp);
        const int zp = static_cast<int>(vz % (D + 2));
        const long long v = vz / (D + 2);
        float c[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) c[e] = 0.f;
        if (xp >= 1 && xp <= W && yp >= 1 && yp <= H && zp >= 1 && zp <= D) {
            const long long vox = (static_cast<long long>(zp - 1) * H + (yp - 1)) * W + (xp - 1);
            if (feat_cl) {          // channels-last producer: the record is already contiguous
                const float4* p = reinterpret_cast<const float4*>(feat + (v * S + vox) * 16);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 q = __ldg(p + e);
                    c[4 * e] = q.x, c[4 * e + 1] = q.y, c[4 * e + 2] = q.z, c[4 * e + 3] = q.w;
                }
            } else {
                const float* p = feat + v * 16 * S + vox;
#pragma unroll
                for (int e = 0; e < 16; ++e) c[e] = __ldg(p + e * S);
            }
        }
        float4* o = feat_pad + i * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = make_float4(c[4 * e], c[4 * e + 1], c[4 * e + 2], c[4 * e + 3]);
    }
    if (i < n_quad) {
        const int xq = static_cast<int>(i % Wq), yq = static_cast<int>((i / Wq) % Hq);
        const long long vz = i / (static_cast<long long>(Wq) * Hq);
        const int z = static_cast<int>(vz % (D + 2)) - 1;
        const long long v = vz / (D + 2);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (z >= 0 && z < D) {
            const float* d = dens + v * S + static_cast<long long>(z) * H * W;
            const int y0 = yq - 1, x0 = xq - 1;
            const bool ya = y0 >= 0, yb = yq < H, xa = x0 >= 0, xb = xq < W;
            if (ya && xa) q.x = __ldg(d + y0 * W + x0);
            if (ya && xb) q.y = __ldg(d + y0 * W + xq);
            if (yb && xa) q.z = __ldg(d + yq * W + x0);
            if (yb && xb) q.w = __ldg(d + yq * W + xq);
        }
        dens_quad[i] = q;
    }
}

}  // namespace forge

extern "C" int forge_pack_volume(const float* feat, int feat_channels_last, const float* dens, float* feat_pad,
                                 float* dens_quad, int V, int D, int H, int W, void* stream) {
    FORGE_RANGE("forge_pack_volume");
    using namespace forge;
    const char* fn = "forge_pack_volume";
    if (!feat || !dens || !feat_pad || !dens_quad) return fail(fn, "null pointer");
    if (V <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(fn, "non-positive size");
    if (V > 65535) return fail(fn, "more than 65535 volumes in one launch");
    if (!aligned16(dens_quad)) return fail(fn, "dens_quad must be 16-byte aligned");
    if (!aligned16(feat_pad) || (feat_channels_last && !aligned16(feat))) return fail(fn, "feat_pad / feat must be 16-byte aligned");
    const long long n_vox = static_cast<long long>(V) * (D + 2) * (H + 2) * (W + 2);
    const long long n_quad = static_cast<long long>(V) * (D + 2) * (H + 1) * (W + 1);
    if ((n_vox + kPackThreads - 1) / kPackThreads > 2147483647LL) return fail(fn, "volume batch too large for one launch");
    pack_volume_kernel<<<static_cast<unsigned>((n_vox + kPackThreads - 1) / kPackThreads), kPackThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, feat_channels_last, dens, reinterpret_cast<float4*>(feat_pad), reinterpret_cast<float4*>(dens_quad), D, H, W, n_vox, n_quad);
    return check_launch(fn);
}

extern "C" int forge_unpack_volume_grad(const float* grad_feat_pad, float* grad_feat, int channels_last, int V, int D,
                                        int H, int W, void* stream) {
    FORGE_RANGE("forge_unpack_volume_grad");
    using namespace forge;
    const char* fn = "forge_unpack_volume_grad";
    if (!grad_feat_pad || !grad_feat) return fail(fn, "null pointer");
    if (V <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(fn, "non-positive size");
    if (V > 65535) return fail(fn, "more than 65535 volumes in one launch");
    const size_t smem = sizeof(float) * 16 * (W + 1);
    if (smem > 48 * 1024) return fail(fn, "volume rows longer than 700 voxels are not supported");
    dim3 grid(D * H, V);
    unpack_grad_kernel<<<grid, kPackThreads, smem, static_cast<cudaStream_t>(stream)>>>(grad_feat_pad, grad_feat,
                                                                                        channels_last, D, H, W);
    return check_launch(fn);
}
