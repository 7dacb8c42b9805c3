// K2: affine resample of feature volumes into the canonical camera frame (forward + backward).
//
// Replaces reference models/rotate.py:127-141: the voxel-centre grid is affine in (w, h, d), so it
// is evaluated in registers and never materialised; volumes are channels-last so the 8 corner
// fetches of an output voxel are contiguous C*4-byte runs (512 B for C = 128) and the output write
// is one contiguous run -- every HBM byte moved is an algorithmic byte.  HBM-bound.
//
// Work mapping: a CTA owns a compact block of output voxels (8x4x4 by default, see forge_rotate_fwd).  Phase 1: one thread per voxel
// evaluates the sample position once and leaves the 8 corner offsets + weights in shared memory
// (out-of-volume corners get weight 0 and a clamped offset, so phase 2 has no predicates).
// Phase 2: one thread per (voxel, channel vector): 4 broadcast LDS.128, 8 coalesced LDG.128,
// 32 FFMA, 1 coalesced STG.128 -- the per-voxel scalar work is not replicated across the 32 channel
// lanes, which is what kept v1 issue-bound at 36 % of HBM peak.
#include <cstdlib>

#include "common.cuh"

namespace forge {

constexpr int kRotThreads = 256;
constexpr int kTileVox = 256;                       // output voxels per CTA = one voxel per thread in phase 1
// output-voxel block per CTA, kTx * kTy * kTz == kTileVox (shape 0 = 8x8x4 is the default)
struct TileShape {
    int tx, ty, tz;
};
__host__ __device__ constexpr TileShape tile_shape(int id) {
    return id == 1 ? TileShape{16, 4, 4} : id == 2 ? TileShape{4, 8, 8} : id == 3 ? TileShape{16, 16, 1} : id == 4 ? TileShape{32, 8, 1}
         : id == 5 ? TileShape{8, 8, 2} : id == 6 ? TileShape{8, 4, 4} : id == 7 ? TileShape{8, 4, 2} : id == 8 ? TileShape{4, 4, 4}
         : TileShape{8, 8, 4};     // 5, 6: half tiles (128 voxels), 7, 8: quarter tiles
}
static_assert(kTileVox == kRotThreads, "phase 1 maps one thread to one voxel");

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

__device__ __forceinline__ float4 vzero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vfma(float& a, float v, float w) { a = fmaf(v, w, a); }
__device__ __forceinline__ void vfma(float4& a, const float4& v, float w) {
    a.x = fmaf(v.x, w, a.x);
    a.y = fmaf(v.y, w, a.y);
    a.z = fmaf(v.z, w, a.z);
    a.w = fmaf(v.w, w, a.w);
}
template <typename VecT>
__device__ __forceinline__ VecT vzero();
template <>
__device__ __forceinline__ float vzero<float>() {
    return 0.f;
}
template <>
__device__ __forceinline__ float4 vzero<float4>() {
    return vzero4();
}

struct RotTile {
    int off[kTileVox][8];      // corner voxel offsets (clamped into the volume)
    float w[kTileVox][8];      // corner weights, 0 for corners outside the volume
    int out[kTileVox];         // linear output voxel index, -1 outside the volume
    unsigned mask[kTileVox];   // in-bounds bits of the 8 corners
    float A[12];
};

// Compact per-voxel records of the lean forward kernel: two broadcast LDS.128 per voxel instead of six (output index, mask, 8
// offsets, 8 weights) -- the kernel is bound by L1 data-pipe wavefronts (ncu: 85 % of peak), and those reads were 6 of its 42 per
// voxel.  cls: 0 = sample outside the source volume, 1 = all 8 corners inside (offsets are base + fixed strides, weights are
// formed in registers with ATen's (wx wy) wz order), 2 = partly outside (the full tables of RotTile are used).
struct RotLean {
    int4 rec[kTileVox];        // {output voxel index or -1, base voxel offset | cls << 30, wx0, wx1}
    float4 wts[kTileVox];      // {wy0, wy1, wz0, wz1}
};

// phase 1 shared by forward and backward; optionally keeps the per-axis weights for d out/d pos
__device__ __forceinline__ void rotate_phase1(RotTile& s, float (*frac)[6], const float* __restrict__ affine, int m,
                                              const float* __restrict__ gx, const float* __restrict__ gy,
                                              const float* __restrict__ gz, float inv_max, int D, int H, int W,
                                              int tx, int ty, int tz, const TileShape sh, RotLean* lean = nullptr) {
    const int kTx = sh.tx, kTy = sh.ty, kTz = sh.tz;
    if (threadIdx.x < 12) s.A[threadIdx.x] = affine[12 * m + threadIdx.x];
    __syncthreads();
    const int v = threadIdx.x;
    const int w = tx * kTx + (v % kTx), h = ty * kTy + ((v / kTx) % kTy), d = tz * kTz + v / (kTx * kTy);
    if (v >= kTx * kTy * kTz || w >= W || h >= H || d >= D) {     // half tiles leave the upper threads without a voxel
        s.out[v] = -1;
        if (lean) lean->rec[v] = make_int4(-1, 0, 0, 0);
    } else {
        s.out[v] = (d * H + h) * W + w;
        const Tri t = rotate_tri(s.A, gx[w], gy[h], gz[d], inv_max, D, H, W);
        s.mask[v] = t.mask;
#pragma unroll
        for (int cn = 0; cn < 8; ++cn) {
            const bool in = (t.mask >> cn) & 1u;
            const int x = min(max(t.x0 + (cn & 1), 0), W - 1), y = min(max(t.y0 + ((cn >> 1) & 1), 0), H - 1),
                      z = min(max(t.z0 + (cn >> 2), 0), D - 1);
            s.off[v][cn] = (z * H + y) * W + x;
            s.w[v][cn] = in ? tri_weight(t, cn) : 0.f;
        }
        if (lean) {
            const int cls = t.mask == 0u ? 0 : (t.mask == 0xffu ? 1 : 2);
            const int base = cls == 1 ? (t.z0 * H + t.y0) * W + t.x0 : 0;
            lean->rec[v] = make_int4(s.out[v], base | (cls << 30), __float_as_int(t.wx0), __float_as_int(t.wx1));
            lean->wts[v] = make_float4(t.wy0, t.wy1, t.wz0, t.wz1);
        }
        if (frac) {
            frac[v][0] = t.wx0;
            frac[v][1] = t.wx1;
            frac[v][2] = t.wy0;
            frac[v][3] = t.wy1;
            frac[v][4] = t.wz0;
            frac[v][5] = t.wz1;
        }
    }
    __syncthreads();
}

// CU = channel vectors per voxel (C/4 for float4, C for float)
template <typename VecT, int kShape, bool kStream>
__global__ void __launch_bounds__(kRotThreads, 4)
rotate_fwd_kernel(const VecT* __restrict__ in, const float* __restrict__ affine, const int* __restrict__ jobs,
                  const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                  float inv_max, VecT* __restrict__ out, int CU, int D, int H, int W, int tiles_x, int tiles_y) {
    __shared__ __align__(16) RotTile s;
    constexpr TileShape sh = tile_shape(kShape);
    constexpr int kTx = sh.tx, kTy = sh.ty, kTz = sh.tz;
    const int m = blockIdx.y;
    const RotJob job = {jobs[3 * m], jobs[3 * m + 1], jobs[3 * m + 2]};
    const int tz = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - tz * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const long long vol = static_cast<long long>(D) * H * W;
    const VecT* src = in + static_cast<long long>(job.src) * vol * CU;
    VecT* dst = out + static_cast<long long>(job.dst) * vol * CU;

    constexpr int kVox = kTx * kTy * kTz;
    if (job.kind == 1) {   // view-0 passthrough (models/rotate.py:141)
        for (int e = threadIdx.x; e < kVox * CU; e += kRotThreads) {
            const int v = e / CU, cu = e - v * CU;
            const int w = tx * kTx + (v % kTx), h = ty * kTy + ((v / kTx) % kTy), d = tz * kTz + v / (kTx * kTy);
            if (w >= W || h >= H || d >= D) continue;
            const long long o = (static_cast<long long>(d) * H + h) * W + w;
            dst[o * CU + cu] = src[o * CU + cu];
        }
        return;
    }
    rotate_phase1(s, nullptr, affine, m, gx, gy, gz, inv_max, D, H, W, tx, ty, tz, sh);

#pragma unroll 2
    for (int e = threadIdx.x; e < kVox * CU; e += kRotThreads) {
        const int v = e / CU, cu = e - v * CU;
        const int o = s.out[v];
        if (o < 0) continue;
        if (s.mask[v] == 0u) {              // the sample lies outside the source volume (zeros padding): nothing to read
            dst[static_cast<long long>(o) * CU + cu] = vzero<VecT>();
            continue;
        }
        const int4 o0 = *reinterpret_cast<const int4*>(&s.off[v][0]), o1 = *reinterpret_cast<const int4*>(&s.off[v][4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&s.w[v][0]), w1 = *reinterpret_cast<const float4*>(&s.w[v][4]);
        const VecT* p = src + cu;
        const VecT v0 = __ldg(p + static_cast<long long>(o0.x) * CU), v1 = __ldg(p + static_cast<long long>(o0.y) * CU),
                   v2 = __ldg(p + static_cast<long long>(o0.z) * CU), v3 = __ldg(p + static_cast<long long>(o0.w) * CU),
                   v4 = __ldg(p + static_cast<long long>(o1.x) * CU), v5 = __ldg(p + static_cast<long long>(o1.y) * CU),
                   v6 = __ldg(p + static_cast<long long>(o1.z) * CU), v7 = __ldg(p + static_cast<long long>(o1.w) * CU);
        VecT acc = vzero<VecT>();      // ATen accumulation order: x fastest, then y, then z
        vfma(acc, v0, w0.x);
        vfma(acc, v1, w0.y);
        vfma(acc, v2, w0.z);
        vfma(acc, v3, w0.w);
        vfma(acc, v4, w1.x);
        vfma(acc, v5, w1.y);
        vfma(acc, v6, w1.z);
        vfma(acc, v7, w1.w);
        if (kStream)
            __stcs(dst + static_cast<long long>(o) * CU + cu, acc);      // written once, never re-read by this kernel
        else
            dst[static_cast<long long>(o) * CU + cu] = acc;
    }
}

// Lean forward for C = 128 (the lifted feature volumes: 32 float4 per voxel = one voxel per warp instruction), round 2.
// Same phases and the same arithmetic as rotate_fwd_kernel; what changes is the instruction stream of phase 2, which the
// generic kernel spends mostly on index arithmetic (an integer division by the runtime channel count per item and eight
// 64-bit multiply-adds: 130 SASS instructions per (voxel, float4), issue slots 54 % busy, ALU pipe 40 %): here a warp owns
// whole voxels, the lane is the channel vector, corner offsets are premultiplied 32-bit float4 indices.
template <int kShape, bool kStream>
__global__ void __launch_bounds__(kRotThreads, 4)       // 5 / 6 CTAs per SM (48 / 40 registers) measured: 0.1157 / 0.1188 vs 0.1147 ms
rotate_fwd_c128_kernel(const float4* __restrict__ in, const float* __restrict__ affine, const int* __restrict__ jobs,
                       const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                       float inv_max, float4* __restrict__ out, int D, int H, int W, int tiles_x, int tiles_y) {
    __shared__ __align__(16) RotTile s;
    constexpr TileShape sh = tile_shape(kShape);
    constexpr int kTx = sh.tx, kTy = sh.ty, kTz = sh.tz, kVox = kTx * kTy * kTz, kCU = 32;
    const int m = blockIdx.y;
    const RotJob job = {jobs[3 * m], jobs[3 * m + 1], jobs[3 * m + 2]};
    const int tz = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - tz * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const long long vol = static_cast<long long>(D) * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* src = in + static_cast<long long>(job.src) * vol * kCU + lane;
    float4* dst = out + static_cast<long long>(job.dst) * vol * kCU + lane;

    if (job.kind == 1) {   // view-0 passthrough (models/rotate.py:141)
        for (int v = warp; v < kVox; v += kRotThreads / 32) {
            const int w = tx * kTx + (v % kTx), h = ty * kTy + ((v / kTx) % kTy), d = tz * kTz + v / (kTx * kTy);
            if (w >= W || h >= H || d >= D) continue;
            const int o = ((d * H + h) * W + w) * kCU;
            dst[o] = src[o];
        }
        return;
    }
    __shared__ __align__(16) RotLean sl;
    rotate_phase1(s, nullptr, affine, m, gx, gy, gz, inv_max, D, H, W, tx, ty, tz, sh, &sl);
    constexpr unsigned long long kVoxBytes = kCU * 16ull;
    const unsigned long long sy_b = static_cast<unsigned long long>(W) * kVoxBytes, sz_b = static_cast<unsigned long long>(H) * sy_b;

#pragma unroll 2
    for (int v = warp; v < kVox; v += kRotThreads / 32) {
        const int4 r = sl.rec[v];
        if (r.x < 0) continue;
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<char*>(dst) + static_cast<unsigned long long>(static_cast<unsigned>(r.x)) * kVoxBytes);
        const unsigned cls = static_cast<unsigned>(r.y) >> 30;
        float4 acc = vzero4();      // ATen accumulation order: x fastest, then y, then z
        if (cls == 1u) {            // all corners inside: base + fixed strides, weights (wx wy) wz formed here
            const float4 q = sl.wts[v];
            const float wx0 = __int_as_float(r.z), wx1 = __int_as_float(r.w);
            const float w00 = __fmul_rn(wx0, q.x), w10 = __fmul_rn(wx1, q.x), w01 = __fmul_rn(wx0, q.y), w11 = __fmul_rn(wx1, q.y);
            const char* p0 = reinterpret_cast<const char*>(src) + static_cast<unsigned long long>(static_cast<unsigned>(r.y) & 0x3fffffffu) * kVoxBytes;
            const char* p1 = p0 + sz_b;
            auto at = [](const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
            const float4 v0 = at(p0), v1 = at(p0 + kVoxBytes), v2 = at(p0 + sy_b), v3 = at(p0 + sy_b + kVoxBytes), v4 = at(p1),
                         v5 = at(p1 + kVoxBytes), v6 = at(p1 + sy_b), v7 = at(p1 + sy_b + kVoxBytes);
            vfma(acc, v0, __fmul_rn(w00, q.z));
            vfma(acc, v1, __fmul_rn(w10, q.z));
            vfma(acc, v2, __fmul_rn(w01, q.z));
            vfma(acc, v3, __fmul_rn(w11, q.z));
            vfma(acc, v4, __fmul_rn(w00, q.w));
            vfma(acc, v5, __fmul_rn(w10, q.w));
            vfma(acc, v6, __fmul_rn(w01, q.w));
            vfma(acc, v7, __fmul_rn(w11, q.w));
        } else if (cls == 2u) {     // partly outside: clamped offsets, masked weights from the full tables
            const int4 o0 = *reinterpret_cast<const int4*>(&s.off[v][0]), o1 = *reinterpret_cast<const int4*>(&s.off[v][4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&s.w[v][0]), w1 = *reinterpret_cast<const float4*>(&s.w[v][4]);
            auto at = [&](int off) {
                return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(src) +
                                                             static_cast<unsigned long long>(static_cast<unsigned>(off)) * kVoxBytes));
            };
            const float4 v0 = at(o0.x), v1 = at(o0.y), v2 = at(o0.z), v3 = at(o0.w), v4 = at(o1.x), v5 = at(o1.y), v6 = at(o1.z),
                         v7 = at(o1.w);
            vfma(acc, v0, w0.x);
            vfma(acc, v1, w0.y);
            vfma(acc, v2, w0.z);
            vfma(acc, v3, w0.w);
            vfma(acc, v4, w1.x);
            vfma(acc, v5, w1.y);
            vfma(acc, v6, w1.z);
            vfma(acc, v7, w1.w);
        }                           // cls == 0: the sample lies outside the source volume (zeros padding): nothing to read
        if (kStream) __stcs(op, acc);       // written once, never re-read by this kernel: leave L2 to the source volumes
        else *op = acc;
    }
}

// ---- backward ------------------------------------------------------------------------------------
__device__ __forceinline__ void vred(float* addr, float v) { atomicAdd(addr, v); }
__device__ __forceinline__ void vred(float4* addr, const float4& v) { red_add_v4(reinterpret_cast<float*>(addr), v); }
__device__ __forceinline__ float vscale(float v, float w) { return v * w; }
__device__ __forceinline__ float4 vscale(const float4& v, float w) { return make_float4(v.x * w, v.y * w, v.z * w, v.w * w); }
__device__ __forceinline__ float vdot(float a, float b) { return a * b; }
__device__ __forceinline__ float vdot(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// grad_vox: scatter g * w_corner with vector REDs (channels-last => contiguous, coalesced).
// grad_affine: d out/d sample-pos needs the corner values; chain through
//   ix = ((s+1) size - 1)/2, s = c * inv_max, c = A . [gx, gy, gz, 1]
// into 12 floats per job, reduced per CTA then one atomicAdd per float.
template <typename VecT>
__global__ void __launch_bounds__(kRotThreads)
rotate_bwd_kernel(const VecT* __restrict__ in, const float* __restrict__ affine, const int* __restrict__ jobs,
                  const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                  float inv_max, const VecT* __restrict__ g_out, VecT* __restrict__ grad_in,
                  float* __restrict__ grad_aff, int CU, int D, int H, int W, int tiles_x, int tiles_y) {
    __shared__ __align__(16) RotTile s;
    constexpr TileShape sh = tile_shape(6);           // half tiles, like the forward kernel
    constexpr int kTx = sh.tx, kTy = sh.ty, kTz = sh.tz, kVox = kTx * kTy * kTz;
    __shared__ float frac[kTileVox][6];
    __shared__ float red[12][kRotThreads / 32];
    const int m = blockIdx.y;
    const RotJob job = {jobs[3 * m], jobs[3 * m + 1], jobs[3 * m + 2]};
    const int tz = blockIdx.x / (tiles_x * tiles_y);
    const int trem = blockIdx.x - tz * tiles_x * tiles_y;
    const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
    const long long vol = static_cast<long long>(D) * H * W;
    const VecT* src = in + static_cast<long long>(job.src) * vol * CU;
    const VecT* go = g_out + static_cast<long long>(job.dst) * vol * CU;
    VecT* gin = grad_in ? grad_in + static_cast<long long>(job.src) * vol * CU : nullptr;

    if (job.kind == 1) {
        if (!gin) return;
        for (int e = threadIdx.x; e < kVox * CU; e += kRotThreads) {
            const int v = e / CU, cu = e - v * CU;
            const int w = tx * kTx + (v % kTx), h = ty * kTy + ((v / kTx) % kTy), d = tz * kTz + v / (kTx * kTy);
            if (w >= W || h >= H || d >= D) continue;
            const long long o = (static_cast<long long>(d) * H + h) * W + w;
            vred(gin + o * CU + cu, go[o * CU + cu]);
        }
        return;
    }
    const bool need_aff = grad_aff != nullptr;
    rotate_phase1(s, frac, affine, m, gx, gy, gz, inv_max, D, H, W, tx, ty, tz, sh);

    const float kx = 0.5f * static_cast<float>(W) * inv_max, ky = 0.5f * static_cast<float>(H) * inv_max,
                kz = 0.5f * static_cast<float>(D) * inv_max;
    float ga[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) ga[e] = 0.f;

    for (int e = threadIdx.x; e < kVox * CU; e += kRotThreads) {
        const int v = e / CU, cu = e - v * CU;
        const int o = s.out[v];
        if (o < 0) continue;
        const unsigned mask = s.mask[v];
        if (mask == 0u) continue;           // sample outside the source volume: no gradient to anything
        const VecT g = go[static_cast<long long>(o) * CU + cu];
        float gix = 0.f, giy = 0.f, giz = 0.f;
#pragma unroll
        for (int cn = 0; cn < 8; ++cn) {
            if ((mask >> cn) & 1u) {     // corners outside the volume are the zero padding: no gradient
                const long long vox = s.off[v][cn];
                if (gin) vred(gin + vox * CU + cu, vscale(g, s.w[v][cn]));
                if (need_aff) {
                    const float q = vdot(g, __ldg(src + vox * CU + cu));
                    const float wx = frac[v][cn & 1], wy = frac[v][2 + ((cn >> 1) & 1)], wz = frac[v][4 + (cn >> 2)];
                    gix = fmaf((cn & 1) ? wy * wz : -(wy * wz), q, gix);
                    giy = fmaf((cn & 2) ? wx * wz : -(wx * wz), q, giy);
                    giz = fmaf((cn & 4) ? wx * wy : -(wx * wy), q, giz);
                }
            }
        }
        if (need_aff) {
            const int w_ = tx * kTx + (v % kTx), h_ = ty * kTy + ((v / kTx) % kTy), d_ = tz * kTz + v / (kTx * kTy);
            const float gxw = gx[w_], gyh = gy[h_], gzd = gz[d_];
            const float rx = gix * kx, ry = giy * ky, rz = giz * kz;
            ga[0] = fmaf(rx, gxw, ga[0]);
            ga[1] = fmaf(rx, gyh, ga[1]);
            ga[2] = fmaf(rx, gzd, ga[2]);
            ga[3] += rx;
            ga[4] = fmaf(ry, gxw, ga[4]);
            ga[5] = fmaf(ry, gyh, ga[5]);
            ga[6] = fmaf(ry, gzd, ga[6]);
            ga[7] += ry;
            ga[8] = fmaf(rz, gxw, ga[8]);
            ga[9] = fmaf(rz, gyh, ga[9]);
            ga[10] = fmaf(rz, gzd, ga[10]);
            ga[11] += rz;
        }
    }
    if (need_aff) {     // uniform per CTA
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            float x = ga[e];
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) x += __shfl_xor_sync(0xffffffffu, x, sft);
            if (lane == 0) red[e][warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < kRotThreads / 32; ++w) x += red[threadIdx.x][w];
            atomicAdd(grad_aff + 12 * m + threadIdx.x, x);
        }
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
    FORGE_RANGE("forge_rotate_fwd");
    using namespace forge;
    const char* fn = "forge_rotate_fwd";
    if (int e = rotate_check(fn, vox_cl, affine12, jobs, gx, gy, gz, grid_coord_max, M, C, D, H, W)) return e;
    if (!out_cl) return fail(fn, "null pointer");
    static const int shape_env = [] {       // tuning knobs (development)
        const char* e = getenv("FORGE_K2_SHAPE");
        return e ? atoi(e) : -1;
    }();
    // default: half tiles (8x4x4 voxels per 256-thread CTA) -- twice the CTAs of a full tile halve the tail of the
    // launch (cfg-2: 4.3 -> 8.6 waves, 61.9 -> 66.0 % of the HBM peak; cfg-4: 71.2 -> 72.4 %); quarter tiles when even
    // those do not fill one wave of the GPU (cfg-1: 0.029 -> 0.016 ms)
    int shape_id = shape_env;
    if (shape_id < 0) {
        const long long half_tiles = static_cast<long long>((W + 7) / 8) * ((H + 3) / 4) * ((D + 3) / 4) * M;
        shape_id = half_tiles < 1184 ? 7 : 6;
    }
    static const bool stream_st = [] {
        const char* e = getenv("FORGE_K2_STREAM");
        return e ? atoi(e) != 0 : false;
    }();
    const TileShape sh = tile_shape(shape_id);
    const int tiles_x = (W + sh.tx - 1) / sh.tx, tiles_y = (H + sh.ty - 1) / sh.ty, tiles_z = (D + sh.tz - 1) / sh.tz;
    dim3 grid(tiles_x * tiles_y * tiles_z, M);
    const float inv_max = 1.0f / grid_coord_max;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool lean = [] {
        const char* e = getenv("FORGE_K2_LEAN");
        return e ? atoi(e) != 0 : true;
    }();
    if (lean && C == 128 && (shape_id == 6 || shape_id == 7 || shape_id == 8 || shape_id == 0) && aligned16(vox_cl) && aligned16(out_cl) &&
        static_cast<long long>(D) * H * W * 32 < 2147483647LL) {
        const float4* in4 = reinterpret_cast<const float4*>(vox_cl);
        float4* out4 = reinterpret_cast<float4*>(out_cl);
#define FORGE_K2_LEAN(SHAPE)                                                                                                    \
    do {                                                                                                                        \
        if (stream_st)                                                                                                          \
            rotate_fwd_c128_kernel<SHAPE, true><<<grid, kRotThreads, 0, st>>>(in4, affine12, jobs, gx, gy, gz, inv_max, out4, D, H, W, \
                                                                             tiles_x, tiles_y);                                 \
        else                                                                                                                    \
            rotate_fwd_c128_kernel<SHAPE, false><<<grid, kRotThreads, 0, st>>>(in4, affine12, jobs, gx, gy, gz, inv_max, out4, D, H, W, \
                                                                              tiles_x, tiles_y);                                \
    } while (0)
        if (shape_id == 6) FORGE_K2_LEAN(6);
        else if (shape_id == 7) FORGE_K2_LEAN(7);
        else if (shape_id == 8) FORGE_K2_LEAN(8);
        else FORGE_K2_LEAN(0);
#undef FORGE_K2_LEAN
    } else if (C % 4 == 0 && aligned16(vox_cl) && aligned16(out_cl)) {
        const float4* in4 = reinterpret_cast<const float4*>(vox_cl);
        float4* out4 = reinterpret_cast<float4*>(out_cl);
#define FORGE_K2_LAUNCH(SHAPE, STREAM)                                                                                  \
    rotate_fwd_kernel<float4, SHAPE, STREAM><<<grid, kRotThreads, 0, st>>>(in4, affine12, jobs, gx, gy, gz, inv_max, out4, \
                                                                          C / 4, D, H, W, tiles_x, tiles_y)
        if (stream_st) {
            switch (shape_id) {
                case 1: FORGE_K2_LAUNCH(1, true); break;
                case 2: FORGE_K2_LAUNCH(2, true); break;
                case 3: FORGE_K2_LAUNCH(3, true); break;
                case 4: FORGE_K2_LAUNCH(4, true); break;
                case 5: FORGE_K2_LAUNCH(5, true); break;
                case 6: FORGE_K2_LAUNCH(6, true); break;
                case 7: FORGE_K2_LAUNCH(7, true); break;
                case 8: FORGE_K2_LAUNCH(8, true); break;
                default: FORGE_K2_LAUNCH(0, true); break;
            }
        } else {
            switch (shape_id) {
                case 1: FORGE_K2_LAUNCH(1, false); break;
                case 2: FORGE_K2_LAUNCH(2, false); break;
                case 3: FORGE_K2_LAUNCH(3, false); break;
                case 4: FORGE_K2_LAUNCH(4, false); break;
                case 5: FORGE_K2_LAUNCH(5, false); break;
                case 6: FORGE_K2_LAUNCH(6, false); break;
                case 7: FORGE_K2_LAUNCH(7, false); break;
                case 8: FORGE_K2_LAUNCH(8, false); break;
                default: FORGE_K2_LAUNCH(0, false); break;
            }
        }
#undef FORGE_K2_LAUNCH
    } else {
        const TileShape s0 = tile_shape(0);
        const int tx0 = (W + s0.tx - 1) / s0.tx, ty0 = (H + s0.ty - 1) / s0.ty, tz0 = (D + s0.tz - 1) / s0.tz;
        rotate_fwd_kernel<float, 0, false><<<dim3(tx0 * ty0 * tz0, M), kRotThreads, 0, st>>>(
            vox_cl, affine12, jobs, gx, gy, gz, inv_max, out_cl, C, D, H, W, tx0, ty0);
    }
    return check_launch(fn);
}

extern "C" int forge_rotate_bwd(const float* vox_cl, const float* affine12, const int* jobs, const float* gx,
                                const float* gy, const float* gz, float grid_coord_max, const float* g_out_cl,
                                float* grad_vox_cl, float* grad_affine12, int M, int C, int D, int H, int W,
                                void* stream) {
    FORGE_RANGE("forge_rotate_bwd");
    using namespace forge;
    const char* fn = "forge_rotate_bwd";
    if (int e = rotate_check(fn, vox_cl, affine12, jobs, gx, gy, gz, grid_coord_max, M, C, D, H, W)) return e;
    if (!g_out_cl) return fail(fn, "null pointer");
    if (!grad_vox_cl && !grad_affine12) return 0;
    const TileShape sh = tile_shape(6);
    const int tiles_x = (W + sh.tx - 1) / sh.tx, tiles_y = (H + sh.ty - 1) / sh.ty, tiles_z = (D + sh.tz - 1) / sh.tz;
    dim3 grid(tiles_x * tiles_y * tiles_z, M);
    const float inv_max = 1.0f / grid_coord_max;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (C % 4 == 0 && aligned16(vox_cl) && aligned16(g_out_cl) && (!grad_vox_cl || aligned16(grad_vox_cl))) {
        rotate_bwd_kernel<float4><<<grid, kRotThreads, 0, st>>>(
            reinterpret_cast<const float4*>(vox_cl), affine12, jobs, gx, gy, gz, inv_max,
            reinterpret_cast<const float4*>(g_out_cl), reinterpret_cast<float4*>(grad_vox_cl), grad_affine12, C / 4, D, H,
            W, tiles_x, tiles_y);
    } else {
        rotate_bwd_kernel<float><<<grid, kRotThreads, 0, st>>>(vox_cl, affine12, jobs, gx, gy, gz, inv_max, g_out_cl,
                                                              grad_vox_cl, grad_affine12, C, D, H, W, tiles_x, tiles_y);
    }
    return check_launch(fn);
}
