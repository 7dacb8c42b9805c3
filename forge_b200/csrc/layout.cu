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
// One CTA per group of kPackRows consecutive padded y-rows of one (volume, z): for every channel those input
// rows are one contiguous run (kPackRows * W floats), read with 128-bit loads, transposed through shared
// memory ([row][x][channel], x-stride 17 / row-stride = 2 mod 32 banks) and written as zero-bordered
// channels-last rows [W+2][16] plus the density-quad rows [W+1][4].  HBM-bound: reads V*17*D*H*W*4 bytes,
// writes ~1.1x that + 4x the density.
constexpr int kPackThreads = 256;

template <int kPackRows>
__global__ void __launch_bounds__(kPackThreads)
pack_volume_kernel(const float* __restrict__ feat, int feat_cl, const float* __restrict__ dens,
                   float* __restrict__ feat_pad, float4* __restrict__ dens_quad, int D, int H, int W, int ygroups) {
    extern __shared__ __align__(16) float sm[];
    const int Wp = W + 2, Hp = H + 2, Wq = W + 1, Hq = H + 1;
    const int TS = W * 17 + 2;                          // floats per transposed row
    float* tile = sm;                                   // [kPackRows][W][17]
    float* drow = sm + kPackRows * TS;                  // [kPackRows + 1][W + 2] density rows y0-1 .. y0+kPackRows-1
    const int v = blockIdx.y;
    const int zp = blockIdx.x / ygroups, yp0 = (blockIdx.x - zp * ygroups) * kPackRows;
    const int z = zp - 1;
    const long long S = static_cast<long long>(D) * H * W;
    const bool zin = (z >= 0 && z < D);
    // interior rows of this group: padded rows yp0 + ry with y = yp0 + ry - 1 in [0, H)
    const int ry_lo = max(0, 1 - yp0), ry_hi = min(kPackRows, H + 1 - yp0);       // [ry_lo, ry_hi)

    if (zin && !feat_cl && ry_hi > ry_lo) {
        const int rows = ry_hi - ry_lo;
        const float* base = feat + static_cast<long long>(v) * 16 * S + (static_cast<long long>(z) * H + (yp0 + ry_lo - 1)) * W;
        if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15u) == 0) {
            const int wq = W >> 2, per_ch = rows * wq;
            for (int u = threadIdx.x; u < 16 * per_ch; u += kPackThreads) {
                const int ch = u / per_ch, rem = u - ch * per_ch;
                const int ry = rem / wq, xq = rem - ry * wq;
                const float4 val = __ldg(reinterpret_cast<const float4*>(base + static_cast<long long>(ch) * S + ry * W) + xq);
                float* t = tile + (ry_lo + ry) * TS + (4 * xq) * 17 + ch;
                t[0] = val.x;
                t[17] = val.y;
                t[34] = val.z;
                t[51] = val.w;
            }
        } else {
            const int per_ch = rows * W;
            for (int u = threadIdx.x; u < 16 * per_ch; u += kPackThreads) {
                const int ch = u / per_ch, rem = u - ch * per_ch;
                const int ry = rem / W, x = rem - ry * W;
                tile[(ry_lo + ry) * TS + x * 17 + ch] = __ldg(base + static_cast<long long>(ch) * S + ry * W + x);
            }
        }
    }
    for (int e = threadIdx.x; e < (kPackRows + 1) * Wp; e += kPackThreads) {
        const int ry = e / Wp, xs = e - ry * Wp;        // xs = x + 1
        const int y = yp0 + ry - 1, x = xs - 1;
        float val = 0.f;
        if (zin && y >= 0 && y < H && x >= 0 && x < W)
            val = __ldg(dens + static_cast<long long>(v) * S + (static_cast<long long>(z) * H + y) * W + x);
        drow[ry * Wp + xs] = val;
    }
    __syncthreads();

    const int nrows = min(kPackRows, Hp - yp0);
    // feature rows: nrows x (Wp * 4) float4, contiguous in feat_pad (consecutive padded rows follow each other)
    float4* out = reinterpret_cast<float4*>(feat_pad + ((static_cast<long long>(v) * (D + 2) + zp) * Hp + yp0) * Wp * 16);
    const int row4 = Wp * 4;
    for (int u = threadIdx.x; u < nrows * row4; u += kPackThreads) {
        const int ry = u / row4, e = u - ry * row4;
        const int xp = e >> 2, c4 = (e & 3) * 4;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (zin && ry >= ry_lo && ry < ry_hi && xp >= 1 && xp <= W) {
            if (feat_cl) {
                const int y = yp0 + ry - 1;
                o = __ldg(reinterpret_cast<const float4*>(feat + (static_cast<long long>(v) * S + (static_cast<long long>(z) * H + y) * W + xp - 1) * 16 + c4));
            } else {
                const float* t = tile + ry * TS + (xp - 1) * 17 + c4;
                o = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        out[u] = o;
    }
    // density quads: row yq = yp covers y = yp - 1 (drow[ry]) and y + 1 = yp (drow[ry + 1]); rows yq <= H
    const int qrows = min(kPackRows, Hq - yp0);
    if (qrows > 0) {
        float4* q = dens_quad + ((static_cast<long long>(v) * (D + 2) + zp) * Hq + yp0) * Wq;
        for (int u = threadIdx.x; u < qrows * Wq; u += kPackThreads) {
            const int ry = u / Wq, xq = u - ry * Wq;
            const float* d0 = drow + ry * Wp, * d1 = d0 + Wp;
            q[u] = make_float4(d0[xq], d0[xq + 1], d1[xq], d1[xq + 1]);
        }
    }
}

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
    static const int rows_env = [] {        // tuning knob (development): padded y-rows per CTA
        const char* e = getenv("FORGE_PACK_ROWS");
        return e ? atoi(e) : 0;
    }();
    // ~37 KB of shared memory per CTA (6 CTAs per SM) is the sweet spot: 8 rows up to W = 64, 4 rows beyond (measured on B200:
    // cfg-2 W = 64: 4 / 8 / 16 rows = 0.046 / 0.044 / 0.058 ms; cfg-4 W = 128: 0.511 / 0.683 / 1.554 ms = 76 / 57 / 25 % of the HBM peak)
    const int rows = (rows_env == 4 || rows_env == 8 || rows_env == 16) ? rows_env : (W > 64 ? 4 : 8);
    const size_t smem = sizeof(float) * (rows * (W * 17 + 2) + (rows + 1) * (W + 2));
    if (smem > 200 * 1024) return fail(fn, "volume rows longer than 360 voxels are not supported");
    const int ygroups = (H + 2 + rows - 1) / rows;
    dim3 grid((D + 2) * ygroups, V);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float4* dq = reinterpret_cast<float4*>(dens_quad);
#define FORGE_PACK(R)                                                                                                  \
    do {                                                                                                               \
        if (int rc = ensure_dynamic_smem(fn, reinterpret_cast<const void*>(pack_volume_kernel<R>), smem)) return rc;   \
        pack_volume_kernel<R><<<grid, kPackThreads, smem, st>>>(feat, feat_channels_last, dens, feat_pad, dq, D, H, W, ygroups); \
    } while (0)
    if (rows == 4) FORGE_PACK(4);
    else if (rows == 16) FORGE_PACK(16);
    else FORGE_PACK(8);
#undef FORGE_PACK
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
