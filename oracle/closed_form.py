"""TEST INFRASTRUCTURE: independent closed-form restatement of the hot path (fp32 or fp64).

Where ``reference_path.py`` walks PyTorch3D's NDC / 4x4-inverse machinery, this file derives the
same quantities directly from the OpenCV pin-hole model, so the two can check each other:

  camera centre        o_w = -R^T t
  pixel ray            d_w(i, j) = R^T [(u - cx)/fx, (v - cy)/fy, 1],  u = j + 1/2, v = i + 1/2
                       (half-resolution image, K already halved: models/volume_render.py:50-51)
  volume local coords  p_l = p_w / ((D - 1)/2 * volume_size / D)   per axis (x<->W, y<->H, z<->D)
  samples              p_k = o + z_k d,  z_k = linspace(min_depth, max_depth, P)
  trilinear fetch      grid_sample(align_corners=True, zeros)      (PyTorch3D VolumeSampler)
  compositing          T_0 = 1, w_k = s_k T_k, T_{k+1} = T_k (1 - s_k)
                       F = sum w_k f_k, O = 1 - T_P, depth = sum w_k z_k   (README.md:26-33)
  origin projection    (fx tx/tz + cx, fy ty/tz + cy)               (models/volume_render.py:77-83)

  rotate               g = linspace(-1,1,n) * (n-1)/2 * volume_size/n;  T = pose_0 pose_1^-1
                       grid = (T [g,1])_{xyz} / max(g); grid_sample(align_corners=False, zeros)
                       (models/rotate.py:48-61, 88, 127-141)
"""
import torch
import torch.nn.functional as F


def camera_rays_local(R, T, K_half, S, vol_dhw, volume_size):
    """-> o_local [N,3], d_local [N,S,S,3] (x,y,z order), all in R's dtype."""
    N = R.shape[0]
    dt, dev = R.dtype, R.device
    D, H, W = vol_dhw
    fx, fy, cx, cy = K_half[:, 0, 0], K_half[:, 1, 1], K_half[:, 0, 2], K_half[:, 1, 2]
    u = torch.arange(S, dtype=dt, device=dev) + 0.5
    dir_x = (u[None, None, :] - cx[:, None, None]) / fx[:, None, None]          # varies with column j
    dir_y = (u[None, :, None] - cy[:, None, None]) / fy[:, None, None]          # varies with row i
    d_cam = torch.stack([dir_x.expand(N, S, S), dir_y.expand(N, S, S), torch.ones(N, S, S, dtype=dt, device=dev)], -1)
    Rt = R.transpose(1, 2)
    d_w = torch.einsum('nab,nijb->nija', Rt, d_cam)
    o_w = -torch.einsum('nab,nb->na', Rt, T)
    scale = torch.tensor([(W - 1) / 2 * (volume_size / D), (H - 1) / 2 * (volume_size / D),
                          (D - 1) / 2 * (volume_size / D)], dtype=dt, device=dev)
    return o_w / scale, d_w / scale


def composite(sigma, feats, zs):
    """sigma [...,P], feats [...,P,C], zs [P] -> feat [...,C], opacity [...], depth [...]."""
    one_minus = 1.0 - sigma
    trans = torch.cumprod(one_minus, dim=-1)
    trans_before = torch.cat([torch.ones_like(trans[..., :1]), trans[..., :-1]], dim=-1)
    w = sigma * trans_before
    feat = (w[..., None] * feats).sum(dim=-2)
    opacity = 1.0 - trans[..., -1]
    depth = (w * zs).sum(dim=-1)
    return feat, opacity, depth, w


def raymarch(R, T, K_half, feature_3d, density_3d, S, n_pts, min_depth, max_depth, volume_size):
    """-> feat [N,S,S,C], opacity [N,S,S], depth [N,S,S] in feature_3d's dtype."""
    N, C, D, H, W = feature_3d.shape
    dt, dev = feature_3d.dtype, feature_3d.device
    o_l, d_l = camera_rays_local(R.to(dt), T.to(dt), K_half.to(dt), S, (D, H, W), volume_size)
    zs = torch.linspace(min_depth, max_depth, n_pts, dtype=torch.float32, device=dev).to(dt)
    pts = o_l[:, None, None, None, :] + zs[None, None, None, :, None] * d_l[:, :, :, None, :]   # [N,S,S,P,3]
    flat = pts.reshape(N, -1, 1, 1, 3)
    sig = F.grid_sample(density_3d, flat, mode='bilinear', padding_mode='zeros', align_corners=True)
    fea = F.grid_sample(feature_3d, flat, mode='bilinear', padding_mode='zeros', align_corners=True)
    sig = sig.reshape(N, S, S, n_pts)
    fea = fea.reshape(N, C, S, S, n_pts).permute(0, 2, 3, 4, 1)
    feat, opacity, depth, _ = composite(sig, fea, zs)
    return feat, opacity, depth


def origin_projection(T, K_half):
    """Pixel (half-res) coordinates of the world origin: models/volume_render.py:77-83."""
    return torch.stack([K_half[:, 0, 0] * T[:, 0] / T[:, 2] + K_half[:, 0, 2],
                        K_half[:, 1, 1] * T[:, 1] / T[:, 2] + K_half[:, 1, 2]], dim=-1)


def rotate_affine(camPoses_cv2):
    """[B,t,4,4] -> [B*(t-1),4,4]: pose_0 @ inverse(pose_1)   (models/rotate.py:88)."""
    B, t = camPoses_cv2.shape[:2]
    p0 = camPoses_cv2[:, :1].expand(B, t - 1, 4, 4).reshape(-1, 4, 4)
    p1 = camPoses_cv2[:, 1:].reshape(-1, 4, 4)
    return p0 @ torch.linalg.inv(p1)


def rotate_axis(n, volume_size, dtype=torch.float32):
    """World coordinate of voxel centre i along one axis, and its max (grid_coord_max)."""
    g = torch.linspace(-1.0, 1.0, n, dtype=torch.float32).to(dtype) * ((n - 1) * (volume_size / n) * 0.5)
    return g, g.max().item()


def rotate(voxels, camPoses_cv2, volume_size=1.0):
    """voxels [B,t,C,D,H,W] -> same shape; view 0 passes through (models/rotate.py:141)."""
    B, t, C, D, H, W = voxels.shape
    dt = voxels.dtype
    A = rotate_affine(camPoses_cv2.to(dt))
    gx, gmax = rotate_axis(W, volume_size, dt)
    gy, _ = rotate_axis(H, volume_size, dt)
    gz, _ = rotate_axis(D, volume_size, dt)
    Z, Y, X = torch.meshgrid(gz, gy, gx, indexing='ij')
    P = torch.stack([X, Y, Z, torch.ones_like(X)], dim=-1).to(voxels.device)           # [D,H,W,4]
    grid = torch.einsum('mab,dhwb->mdhwa', A[:, :3, :], P) / gmax
    out = F.grid_sample(voxels[:, 1:].reshape(B * (t - 1), C, D, H, W), grid, mode='bilinear',
                        padding_mode='zeros', align_corners=False)
    return torch.cat([voxels[:, :1], out.reshape(B, t - 1, C, D, H, W)], dim=1)
