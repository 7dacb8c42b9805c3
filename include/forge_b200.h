/* forge_b200.h -- C ABI of the B200-native FORGE render / rotate hot path.
 *
 * The reference (UT-Austin-RPL/FORGE) is pure Python and has no FFI layer; its hot path is the
 * library-kernel sequence that models/volume_render.py and models/rotate.py trigger through
 * PyTorch3D / ATen.  Each entry point below replaces one such sequence (cited per function,
 * paths relative to the reference root).  The Python modules in forge_b200/models/ bind these
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - return 0 on success, non-zero on error; forge_last_error() gives the thread-local message
 *   - the caller owns every buffer (device memory unless stated), nothing is allocated, no
 *     synchronisation, no stream creation: work is enqueued on `stream` (a cudaStream_t)
 *   - device pointers must be 16-byte aligned; all tensors are dense fp32 / int32
 *   - "channels-last" volume = [V][D][H][W][C] (C fastest); images are [N][S_h][S_w][C]
 *   - built for sm_100a only; there is no CPU fallback
 */
#ifndef FORGE_B200_H
#define FORGE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FORGE_ABI_VERSION 18
#define FORGE_FEAT_CHANNELS 16 /* render feature channels (models/encoder.py:16-22 -> 16) */

int forge_abi_version(void);
const char* forge_last_error(void);

/* ---- layout ------------------------------------------------------------------------------
 * [n][C][S] <-> [n][S][C] transposes (S = D*H*W).  Replaces the strided NCDHW gathers of ATen's
 * grid_sampler_3d (reached from models/rotate.py:137 and PyTorch3D's VolumeSampler via
 * models/volume_render.py:63) by one coalesced re-layout per DISTINCT volume. */
int forge_ncs_to_nsc(const float* src, float* dst, int n, int C, long long S, void* stream);
int forge_nsc_to_ncs(const float* src, float* dst, int n, int C, long long S, void* stream);

/* ---- render-volume packing ------------------------------------------------------------------
 * Builds the two device layouts K1 reads, once per DISTINCT volume (the reference materialises one
 * copy per rendered view, models/model.py:138-139, and ATen's sampler then gathers 17 channels that
 * lie 1 MiB apart):
 *   feat_pad  [V][D+2][H+2][W+2][16]  channels-last, one-voxel zero border (= zeros padding)
 *   dens_quad [V][D+2][H+1][W+1][4]   (d(z,y,x), d(z,y,x+1), d(z,y+1,x), d(z,y+1,x+1)); index
 *                                     (z+1, y+1, x+1) for z in [-1,D], y in [-1,H-1], x in [-1,W-1]
 * feat is [V][16][D][H][W] (feat_channels_last = 0) or [V][D][H][W][16] (= 1); dens is [V][D][H][W].
 * forge_unpack_volume_grad maps a gradient in feat_pad layout back to feat's layout. */
int forge_pack_volume(const float* feat, int feat_channels_last, const float* dens, float* feat_pad,
                      float* dens_quad, int V, int D, int H, int W, void* stream);
int forge_unpack_volume_grad(const float* grad_feat_pad, float* grad_feat, int channels_last, int V,
                             int D, int H, int W, void* stream);

/* ---- K1: fused volume raymarcher ----------------------------------------------------------
 * Replaces models/volume_render.py:53-63 = PyTorch3D cameras_from_opencv_projection +
 * NDCGridRaysampler + VolumeSampler (2x grid_sample, align_corners=True, zeros padding) +
 * EmissionAbsorptionRaymarcher with the depth patch of README.md:26-33.
 *
 *   sample point   p(n,i,j,k) = o_n + zs[k] * (M_n . [j+0.5, i+0.5, 1]^T)      (volume-local coords)
 *   cam12[n]       = { o_n[3], M_n[3][3] row-major }
 *   compositing    T_0 = 1; w_k = s_k T_k; T_{k+1} = T_k (1 - s_k);
 *                  feat = sum w_k f_k; sil = 1 - T_P; depth = sum w_k zs[k]
 *
 *   feat_pad, dens_quad: see forge_pack_volume;  view2vol [N] (volume index of each view)
 *   out_feat [N][S_h][S_w][16]  out_sil [N][S_h][S_w]  out_depth [N][S_h][S_w] or NULL
 */
int forge_raymarch_fwd(const float* feat_pad, const float* dens_quad, const int* view2vol,
                       const float* cam12, const float* zs, float* out_feat, float* out_sil,
                       float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                       void* stream);

/* The two formulations of the forward kernel behind forge_raymarch_fwd, callable directly (same arguments, same
 * results to fp32 summation order): `_gather` = direct 256-bit L1 gathers of the 8 corners (raymarch.cu),
 * `_tma` = voxel bricks of a 16x8 pixel tile x k-slab staged in shared memory by TMA bulk copies behind an
 * mbarrier ring, corners read with conflict-free LDS.128 (raymarch_tma.cu).  forge_raymarch_fwd picks one
 * (environment override FORGE_K1_IMPL=gather|tma, read once). */
int forge_raymarch_fwd_gather(const float* feat_pad, const float* dens_quad, const int* view2vol,
                              const float* cam12, const float* zs, float* out_feat, float* out_sil,
                              float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                              void* stream);
int forge_raymarch_fwd_tma(const float* feat_pad, const float* dens_quad, const int* view2vol,
                           const float* cam12, const float* zs, float* out_feat, float* out_sil,
                           float* out_depth, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                           void* stream);

/* Backward of forge_raymarch_fwd on the same packed inputs.  g_* are the upstream gradients
 * (g_depth may be NULL).  grad_feat_pad (feat_pad layout) and grad_dens_pad [V][D+2][H+2][W+2]
 * (zero-bordered, voxel (z,y,x) at (z+1,y+1,x+1)) are ACCUMULATED into (caller zeroes them);
 * either may be NULL to skip that gradient (pose-only refinement, reference
 * kubric_eval.py:450-504).  grad_cam12 [N][12] is accumulated into (caller zeroes), may be NULL.
 * Replaces the autograd graph PyTorch builds through the sequence named above. */
int forge_raymarch_bwd(const float* feat_pad, const float* dens_quad, const int* view2vol,
                       const float* cam12, const float* zs, const float* g_feat, const float* g_sil,
                       const float* g_depth, float* grad_feat_pad, float* grad_dens_pad, float* grad_cam12,
                       float* workspace, int N, int V, int D, int H, int W, int S_h, int S_w, int P,
                       void* stream);
/* Bytes of caller-owned, 16-byte aligned scratch forge_raymarch_bwd needs (per-ray sigma_k, a_k, T_k between its two
 * passes + the density gradient in dens_quad layout before it is folded into grad_dens_pad; uninitialised is fine). */
long long forge_raymarch_bwd_workspace(int N, int V, int D, int H, int W, int S_h, int S_w, int P);

/* ---- fused decoder ---------------------------------------------------------------------------
 * relu(conv_rgb(x)) of models/volume_render.py:29-37,73 for inference (BatchNorm in eval mode):
 * ConvTranspose2d(16,16,k6,s2,p2)+BN+LeakyReLU -> Conv2d(16,8,k5,p2)+BN+LeakyReLU -> Conv2d(8,3,k5,p2)
 * -> ReLU in one kernel; the two intermediate images stay in shared memory.
 *   x_nhwc [N][S_h][S_w][16] (the raymarcher's out_feat)  ->  rgb_nchw [N][3][2 S_h][2 S_w]
 *   wpack: forge_decoder_wpack_floats() floats = BN-folded weights
 *     W1[py][px][ty][tx][ci][co] = Wt[ci][co][py+2ty][px+2tx] * s1[co]      (4*9*16*16)
 *     W2[ky][kx][ci][co]         = W2[co][ci][ky][kx] * s2[co]              (25*16*8)
 *     W3[ky][kx][ci][4]          = W3[co][ci][ky][kx], co = 3 zero          (25*8*4)
 *     b1[16] = (bt - mean1) s1 + beta1,  b2[8] likewise,  b3[4]             s = gamma / sqrt(var + eps) */
int forge_decoder_wpack_floats(void);
int forge_decoder_fwd(const float* x_nhwc, const float* wpack, float* rgb_nchw, unsigned* sign_masks, int N, int S_h,
                      int S_w, void* stream);

/* Both forward decoders optionally (sign_masks != NULL) emit one uint32 per output pixel [N][2 S_h][2 S_w] with the
 * signs of the three pre-activations (bits 0-15 layer 1, 16-23 layer 2, 24-26 rgb; 1 = positive): all the backward
 * pass needs from the forward one, since the decoder is piece-wise linear.
 *
 * forge_decoder_bwd_data: g_x = (d relu(conv_rgb(x)) / d x)^T g_rgb in one fp32 kernel -- the pose-only backward of
 * models/volume_render.py:73 when the decoder weights are constants (test-time pose optimisation,
 * kubric_eval.py:450-504).  g_rgb [N][3][2S][2S], masks from the forward call, g_x [N][S][S][16] (written).
 *   wpack_bwd (forge_decoder_bwd_wpack_floats() floats, 16-byte aligned), BN scales folded like the forward pack:
 *     W3b[ky][kx][co < 3][c < 8]    = W3[co][c][4-ky][4-kx]
 *     W2b[ky][kx][co < 8][ci < 16]  = W2[co][ci][4-ky][4-kx] * s2[co]
 *     Wd[u][v][co < 16][ci < 16]    = Wt[ci][co][u][v] * s1[co] */
int forge_decoder_bwd_wpack_floats(void);
int forge_decoder_bwd_data(const float* g_rgb_nchw, const unsigned* sign_masks, const float* wpack_bwd, float* g_x_nhwc,
                           int N, int S_h, int S_w, void* stream);

/* ---- tensor-core decoder (bf16 operands, fp32 accumulate; tcgen05.mma + TMEM) ---------------
 * Same function as forge_decoder_fwd -- relu(conv_rgb(x)), models/volume_render.py:29-37,73, eval-mode
 * BN -- computed as implicit GEMMs on the 5th-generation tensor cores: the "bf16 decoder" of
 * BASELINE.json configs[2].  Inputs/outputs stay fp32 (x is rounded to bf16 on load, the two hidden
 * activations are rounded to bf16 after BN + LeakyReLU); deviation from the fp32 module ~3e-3.
 *   wpack (forge_decoder_tc_wpack_bytes() bytes, 16-byte aligned) = bf16 UMMA B operands built from 8 x 8
 *   core matrices blk[n % 8][k % 8] (128 B, un-swizzled K-major), then the fp32 epilogue constants:
 *     layer 1: 9 tiles (input shift a*3+b) of [k/8 = 2][n = 64][k%8]: n = (py*2+px)*16 + co, k = ci,
 *              value Wt[ci][co][py+4-2a][px+4-2b] * s1[co]
 *     layer 2: 10 strips (ky*2 + ci/8) of 13 blocks [8 x zero, W(kx=4), .., W(kx=0)], W(kx)[co][ci%8] =
 *              W2[co][ci][ky][kx] * s2[co]; then 8 zero blocks.  The B operand of source column j (0..11) is the 8-block
 *              window starting at block 12-j: block delta holds W(kx = j - delta) or zero.
 *     layer 3: 5 strips (ky) of the same shape with W(kx)[co < 3][ci] = W3[co][ci][ky][kx], one all-zero strip
 *              (the phantom row ky = 5), then 8 zero blocks.
 *     fp32 b1[16] b2[8] b3[4] (+ padding to 256 B): y = acc + b per layer; s = gamma / sqrt(var + eps),
 *     b1 = (bias_t - mean1) s1 + beta1, b2 likewise, b3 = the last conv's bias.
 *   max_ctas: 0 = two persistent CTAs per SM, otherwise an upper bound on the grid. */
int forge_decoder_tc_wpack_bytes(void);
int forge_decoder_tc_fwd(const float* x_nhwc, const void* wpack, float* rgb_nchw, unsigned* sign_masks, int N, int S_h,
                         int S_w, int max_ctas, void* stream);

/* Test hook: ONE tcgen05.mma (M=128, N=16, K=16, bf16 x bf16 -> fp32, both operands K-major without
 * swizzle) over a caller-built shared-memory image; pins the descriptor semantics the decoder relies on:
 *   A[m][k] at a_off + (k/8) a_lbo + (m/8) a_sbo + (m%8) 16 + (k%8) 2,   B[n][k] likewise,   D = A B^T. */
int forge_umma_probe(const void* image, int image_bytes, unsigned a_off, unsigned a_lbo, unsigned a_sbo,
                     unsigned b_off, unsigned b_lbo, unsigned b_sbo, float* out_128x16, void* stream);

/* ---- K2: affine feature-volume resample ---------------------------------------------------
 * Replaces models/rotate.py:127-141: materialised homogeneous grid, matmul with T^T, divide by
 * grid_coord_max, F.grid_sample(bilinear, zeros, align_corners=False), and the torch.cat that
 * passes view 0 through.  One launch performs M jobs:
 *
 *   jobs[m] = { src volume, dst volume, kind }   kind 0: resample with affine12[m], kind 1: copy
 *   sample  g = (affine12[m] . [gx[w], gy[h], gz[d], 1]^T) * (1 / grid_coord_max)
 *
 *   vox_cl [n_src][D][H][W][C] -> out_cl [n_dst][D][H][W][C]; gx [W], gy [H], gz [D] are the
 *   world coordinates of voxel centres (models/rotate.py:48-52).  The dst index lets the caller
 *   fold the distance-sorted view permutation (models/model.py:152-168) into the resample. */
int forge_rotate_fwd(const float* vox_cl, const float* affine12, const int* jobs, const float* gx,
                     const float* gy, const float* gz, float grid_coord_max, float* out_cl,
                     int M, int C, int D, int H, int W, void* stream);

/* Backward of forge_rotate_fwd.  grad_vox_cl [n_src][D][H][W][C] is accumulated into (caller
 * zeroes) or NULL; grad_affine12 [M][12] is accumulated into (caller zeroes) or NULL (needs
 * vox_cl).  Copy jobs add g_out into grad_vox_cl. */
int forge_rotate_bwd(const float* vox_cl, const float* affine12, const int* jobs, const float* gx,
                     const float* gy, const float* gz, float grid_coord_max, const float* g_out_cl,
                     float* grad_vox_cl, float* grad_affine12, int M, int C, int D, int H, int W,
                     void* stream);

/* ---- camera / pose algebra (host-glue kernels, one launch each) -----------------------------
 * forge_camera_prep_fwd replaces cameras_from_opencv_projection + the NDC un-projection + Volumes.world_to_local
 * that models/volume_render.py:53-61 runs through PyTorch3D, and transform_points_screen(0) of :77-83 / :97-103:
 *   R [N][3][3], T [N][3], K_half [N][3][3] (OpenCV, intrinsics already halved; only fx, fy, cx, cy are read)
 *   s = ((W-1)/2, (H-1)/2, (D-1)/2) * volume_size / D        (sx, sy, sz)
 *   cam12[n] = { o = -R^T t / s,  M = diag(1/s) R^T K^-1 }     sample point of pixel (i, j) at depth z: o + z M (j+.5, i+.5, 1)^T
 *   origin_proj[n] = (fx tx / tz + cx, fy ty / tz + cy), |tz| clamped to eps with its sign kept; may be NULL.
 * forge_camera_prep_bwd: gradients of both outputs w.r.t. R, T, K_half (each output / gradient pointer may be NULL;
 * grad_K is written densely, zeros outside fx, fy, cx, cy). */
int forge_camera_prep_fwd(const float* R, const float* T, const float* K_half, int N, float sx, float sy, float sz,
                          float eps, float* cam12, float* origin_proj, void* stream);
int forge_camera_prep_bwd(const float* R, const float* T, const float* K_half, int N, float sx, float sy, float sz,
                          float eps, const float* g_cam12, const float* g_origin_proj, float* grad_R, float* grad_T,
                          float* grad_K, void* stream);

/* forge_upsample2x_fwd/bwd: F.upsample(mode='bilinear') (align_corners=False) of the silhouette and depth maps to
 * exactly twice the size (models/volume_render.py:69,74); one launch for one or two maps (src1 / dst1 may be NULL).
 *   src [M][S_h][S_w] -> dst [M][2 S_h][2 S_w];  bwd writes (does not accumulate) g_src from g_dst. */
int forge_upsample2x_fwd(const float* src0, const float* src1, float* dst0, float* dst1, int M, int S_h, int S_w,
                         void* stream);
int forge_upsample2x_bwd(const float* g_dst0, const float* g_dst1, float* g_src0, float* g_src1, int M, int S_h,
                         int S_w, void* stream);

/* forge_pose_affine_fwd replaces Rotate_world.get_transformation (models/rotate.py:64-89: repeat, reshape,
 * torch.inverse, matmul) and the identity row of the passthrough view:
 *   poses [B][t][4][4] camera-to-world; affine12[b*t + v] = (poses[b][0] @ inverse(poses[b][v]))[:3, :] for v >= 1,
 *   identity for v = 0; pose_inv [B*t][4][4] (optional) receives inverse(poses) for the backward pass;
 *   *singular_flag (optional, device int) is set to 1 when a pose is singular (its outputs are NaN). */
int forge_pose_affine_fwd(const float* poses, int B, int t, float* affine12, float* pose_inv, int* singular_flag,
                          void* stream);

/* ---- ConvGRU cell, elementwise stages (models/fusion.py:21-35) -----------------------------------
 * The two convolutions of a cell stay cuDNN; the chain between them is two kernels forward and two backward:
 *   gate:  xhr = cat(x, h * sigmoid(g[:, C:]))                      g = conv_gate(cat(x, h))  [B][2C][S]
 *   out:   h'  = h (1 - u) + tanh(o) u,  u = sigmoid(g[:, :C])      o = out_gate(xhr)         [B][C][S]
 * channels_last = 0: tensors are [B][CC][S]; 1: [B][S][CC] (channels_last_3d).  g, o and their gradients are fp32
 * (bf16 flag 0) or bf16 (flag 1, autocast); h, x, xhr, h' and their gradients are fp32.  h and x may have their own
 * batch stride (elements); every other tensor is dense.  S = D*H*W, B * 2C * S < 2^31.
 * Backward: gate_bwd writes d_g[:, C:], d_h (= contribution through h*r), d_x and ZEROES d_g[:, :C]; out_bwd writes
 * d_o, d_g[:, :C], d_h (= contribution through the lerp) and ZEROES d_g[:, C:] -- the caller adds the two d_g / d_h. */
int forge_gru_gate_fwd(const void* g, int g_bf16, const float* h, long long h_batch_stride, const float* x,
                       long long x_batch_stride, float* xhr, int channels_last, int B, int C, int S, void* stream);
int forge_gru_gate_bwd(const float* d_xhr, const void* g, int g_bf16, const float* h, long long h_batch_stride, void* d_g,
                       float* d_h, float* d_x, int channels_last, int B, int C, int S, void* stream);
int forge_gru_out_fwd(const void* o, const void* g, int og_bf16, const float* h, long long h_batch_stride, float* h_new,
                      int channels_last, int B, int C, int S, void* stream);
int forge_gru_out_bwd(const float* d_h_new, const void* o, const void* g, int og_bf16, const float* h,
                      long long h_batch_stride, void* d_o, void* d_g, float* d_h, int channels_last, int B, int C, int S,
                      void* stream);

/* ---- test hook: the index path of both samplers -------------------------------------------
 * For each normalised point pts[m] = (x, y, z) returns the base voxel (floor) index base[m] =
 * (ix0, iy0, iz0) and an 8-bit in-bounds mask (bit = dz*4 + dy*2 + dx) computed by exactly the
 * device functions K1 (align_corners=1) and K2 (align_corners=0) use.  Bit-exact contract
 * against ATen's GridSampler.h unnormalize + floor. */
int forge_sample_points(const float* pts, int M, int D, int H, int W, int align_corners, int* base,
                        unsigned char* mask, void* stream);

/* ---- 3x3x3 convolution on the tcgen05 tensor cores with fused ConvGRU epilogues ----------------------------------
 * Replaces the cuDNN convolutions and the elementwise chain of models/fusion.py:18-35 (ConvGRUCell_3D) and :61-68
 * (fusion_conv) on the inference / pose-refinement path.  bf16 operands, fp32 accumulation (TMEM), fp32 state.
 *
 *   x  [B][D][H][W][Cx] bf16 channels-last, batch stride x_batch_stride ELEMENTS (a view of a [B][t][...] sequence is fine)
 *   h2 optional second source [B][D][H][W][Ch] bf16 (the convolution runs over cat(x, h2) without materialising it)
 *   wpack [27][(Cx + Ch) / 64][Cout][64] bf16: w[co][ci][dz][dy][dx] at [(dz*3+dy)*3+dx][ci / 64][co][ci % 64]
 *   D, H multiples of 4, W a multiple of 8; Cx, Ch multiples of 64; Cout 128 or 256
 *   mode 0 (plain): y = act(acc * scale + shift), scale nullable (= 1), act = LeakyReLU(0.01) when lrelu != 0;
 *                   out_f32 / out_bf16 [B][D][H][W][Cout] (either may be NULL)
 *   mode 1 (gate):  Cout = 2C; u = sigmoid(acc[:, :C] + shift), r = sigmoid(acc[:, C:] + shift); out_f32 = u [..][C],
 *                   out_bf16 = h_state * r [..][C]           (h_state fp32 [B][D][H][W][C])
 *   mode 2 (out):   Cout = C; c = tanh(acc + shift); h' = h_state (1 - u_in) + c u_in; out_f32 = h', out_bf16 = h' (nullable),
 *                   out_norm = h' * scale + norm_shift (nullable; fusion_norm in eval mode)
 *   aux (nullable, fp32 [..][C]): gate mode stores r, out mode stores c -- what the backward pass of the cell needs
 *   mode 3 (shuffle): ConvTranspose3d(k 4, stride 2, pad 1) as a 27-tap GEMM with Cout = 8 parity classes x 32 channels
 *                   (models/encoder.py:17, :26): columns [32 q, 32 q + 32) of input voxel (z, y, x) go, after scale / shift /
 *                   LeakyReLU, to out_bf16[n][2z + qz][2y + qy][2x + qx][out_off .. out_off + 32) in rows of out_pitch channels
 *   mode 4 (heads): Cout = 32: columns [0, 16) * scale + shift -> out_f32 [..][16]; columns [16, 24): + LeakyReLU -> aux [..][8]
 *                   (the second convolutions of features_head / density_head over a 64-channel stem tensor, encoder.py:20-21, :29-31)
 *   Cout = 32 is allowed in modes 0 and 4 only
 *   max_ctas: 0 = one persistent CTA per SM */
int forge_conv3d_tc(const void* x, long long x_batch_stride, int Cx, const void* h2, long long h_batch_stride, int Ch,
                    const void* wpack, int mode, int lrelu, const float* scale, const float* shift,
                    const float* norm_shift, const float* h_state, const float* u_in, float* out_f32, void* out_bf16,
                    float* out_norm, float* aux, int out_pitch, int out_off, int B, int D, int H, int W, int Cout, int max_ctas,
                    void* stream);

/* Conv3d(8 -> 1, 3x3x3, pad 1) + ReLU on channels-last rows of 8 floats (last layer of density_head, models/encoder.py:32-33);
 * w [27][8] = weight[0][ci][dz][dy][dx] at [(dz*3+dy)*3+dx][ci]; y [B][D][H][W] */
int forge_conv3d_c8_to_1_relu(const float* x, const float* w, float bias, float* y, int B, int D, int H, int W, void* stream);

/* Elementwise stages of the backward pass of the tensor-core ConvGRU cell (constant weights), dense channels-last rows
 * [N = B D H W][C] fp32 unless noted; between them run the two transposed convolutions (forge_conv3d_tc, plain mode):
 *   stage 0: a0 dh', a1 u, a2 c, a3 h                       -> o_bf16 d_o [N][C], o0 dgu, o1 dh_dir
 *   stage 1: a0 g1 [N][2C] = [dx_o | d(h r)], a1 r, a2 h, a3 dgu, a4 dh_dir -> o_bf16 dg [N][2C], o0 dh_acc
 *   stage 2: a0 g1 [N][2C], a1 g2 [N][2C] = [dx_g | dh_g], a2 dh_acc        -> o0 dx, o1 dh
 * (derivation: reference models/fusion.py:29-35) */
int forge_gru_tc_bwd(int stage, const float* a0, const float* a1, const float* a2, const float* a3, const float* a4,
                     void* o_bf16, float* o0, float* o1, long long N, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FORGE_B200_H */
