"""Seeded synthetic inputs of the shapes FORGE feeds the render / rotate path.

No dataset is reachable offline; these follow the reference's geometry: canonical extrinsics
``[I | (0, 0, camera_z)]`` for view 0 (reference dataset/kubric.py:100-103), intrinsics
``K = [[1.38888*img, 0, img/2], [0, 1.38888*img, img/2], [0, 0, 1]]`` (reference demo.py:39-41),
``min/max_depth = 0.5/2.0``, ``volume_size = 1.0`` (config/kubric/joint_pose_2d3d.yaml:27-32).
"""
import math
from types import SimpleNamespace

import torch


def make_config(img_size=256, n_pts_per_ray=64, min_depth=0.5, max_depth=2.0, volume_size=1.0, k_size=5,
                camera_z=1.5, dataset_name='kubric', padding_mode='zeros', use_gt_pose=True,
                canonicalize=True, parameter='joint'):
    """A config object with the keys the hot path reads."""
    return SimpleNamespace(
        dataset=SimpleNamespace(img_size=img_size, name=dataset_name),
        render=SimpleNamespace(n_pts_per_ray=n_pts_per_ray, min_depth=min_depth, max_depth=max_depth,
                               volume_size=volume_size, k_size=k_size, camera_z=camera_z),
        network=SimpleNamespace(padding_mode=padding_mode),
        train=SimpleNamespace(use_gt_pose=use_gt_pose, canonicalize=canonicalize, parameter=parameter),
    )


def look_at_extrinsics(cam_pos):
    """OpenCV extrinsics (x right, y down, z forward) of cameras at ``cam_pos`` [n,3] looking at 0."""
    z = -cam_pos / cam_pos.norm(dim=-1, keepdim=True)
    down = torch.tensor([0.0, 1.0, 0.0], dtype=cam_pos.dtype).expand_as(z)
    x = torch.cross(down, z, dim=-1)
    x = x / x.norm(dim=-1, keepdim=True)
    y = torch.cross(z, x, dim=-1)
    R_c2w = torch.stack([x, y, z], dim=-1)           # columns are the camera axes in world coords
    R = R_c2w.transpose(1, 2)
    T = -torch.einsum('nab,nb->na', R, cam_pos)
    return R, T


def ring_cameras(n_views, camera_z=1.5, seed=0, jitter=True):
    """View 0 canonical; the others on a ring around the object with seeded jitter
    (+-10 deg elevation, +-5 % distance).  -> R [n,3,3], T [n,3], pose_c2w [n,4,4] (fp32)."""
    g = torch.Generator().manual_seed(seed)
    az = torch.arange(n_views, dtype=torch.float64) * (2 * math.pi / max(n_views, 1))
    el = torch.zeros(n_views, dtype=torch.float64)
    dist = torch.full((n_views,), float(camera_z), dtype=torch.float64)
    if jitter and n_views > 1:
        az[1:] += (torch.rand(n_views - 1, generator=g, dtype=torch.float64) - 0.5) * 0.3
        el[1:] = (torch.rand(n_views - 1, generator=g, dtype=torch.float64) - 0.5) * (20 * math.pi / 180)
        dist[1:] *= 1 + (torch.rand(n_views - 1, generator=g, dtype=torch.float64) - 0.5) * 0.1
    pos = torch.stack([-dist * torch.cos(el) * torch.sin(az), dist * torch.sin(el),
                       -dist * torch.cos(el) * torch.cos(az)], dim=-1)
    R, T = look_at_extrinsics(pos)
    E = torch.eye(4, dtype=torch.float64).repeat(n_views, 1, 1)
    E[:, :3, :3] = R
    E[:, :3, 3] = T
    pose = torch.linalg.inv(E)
    return R.float(), T.float(), pose.float()


def intrinsics(n, img_size):
    K = torch.tensor([[1.38888 * img_size, 0.0, img_size / 2.0],
                      [0.0, 1.38888 * img_size, img_size / 2.0],
                      [0.0, 0.0, 1.0]], dtype=torch.float32)
    return K.repeat(n, 1, 1)


def render_inputs(n_objects, n_views, img_size, vol, seed=0, dense=False, device='cpu'):
    """Distinct volumes + per-view cameras for ``n_objects`` x ``n_views`` renders.

    -> dict(feat [b,16,D,D,D], dens [b,1,D,D,D], R [b*v,3,3], T [b*v,3], K [b*v,3,3], view2vol [b*v])
    ``dense=True`` uses relu(randn)*1.5 densities (exercises sigma > 1)."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(n_objects, 16, vol, vol, vol, generator=g)
    if dense:
        dens = torch.relu(torch.randn(n_objects, 1, vol, vol, vol, generator=g)) * 1.5
    else:
        dens = 0.3 * torch.rand(n_objects, 1, vol, vol, vol, generator=g)
    Rs, Ts = [], []
    for b in range(n_objects):
        R, T, _ = ring_cameras(n_views, seed=seed * 1000 + b)
        Rs.append(R)
        Ts.append(T)
    out = dict(feat=feat, dens=dens, R=torch.cat(Rs), T=torch.cat(Ts), K=intrinsics(n_objects * n_views, img_size),
               view2vol=torch.arange(n_objects).repeat_interleave(n_views).int())
    return {k: v.to(device) for k, v in out.items()}


def rotate_inputs(n_objects, n_views, channels, grid, seed=0, device='cpu'):
    """voxels [b,t,C,n,n,n] and camera-to-world poses [b,t,4,4] (view 0 canonical)."""
    g = torch.Generator().manual_seed(seed)
    vox = torch.randn(n_objects, n_views, channels, grid, grid, grid, generator=g)
    poses = torch.stack([ring_cameras(n_views, seed=seed * 1000 + b)[2] for b in range(n_objects)])
    return vox.to(device), poses.to(device)


def kubric_batch(n_objects, n_views_all=10, img_size=256, camera_z=1.5, seed=0):
    """A batch dict of the shape reference dataset/kubric.py:390-402 produces (random images, ring cameras
    canonicalised so that view 0 has extrinsics [I | (0,0,camera_z)])."""
    g = torch.Generator().manual_seed(seed)
    E, P = [], []
    for b in range(n_objects):
        R, T, pose = ring_cameras(n_views_all, camera_z=camera_z, seed=seed * 1000 + b)
        e = torch.eye(4).repeat(n_views_all, 1, 1)
        e[:, :3, :3], e[:, :3, 3] = R, T
        E.append(e)
        P.append(pose)
    E, P = torch.stack(E), torch.stack(P)
    rel = torch.linalg.inv(P[:, :1]) @ P                      # pose of view i in view 0's frame
    every2 = torch.linalg.inv(P[:, :-1]) @ P[:, 1:]
    return {
        'images': torch.rand(n_objects, n_views_all, 3, img_size, img_size, generator=g),
        'fg_probabilities': (torch.rand(n_objects, n_views_all, 1, img_size, img_size, generator=g) > 0.5).float(),
        'K_cv2': intrinsics(n_objects * n_views_all, img_size).reshape(n_objects, n_views_all, 3, 3),
        'cam_extrinsics_cv2': E.clone(), 'cam_poses_cv2': P.clone(),
        'cam_extrinsics_cv2_canonicalized': E, 'cam_poses_cv2_canonicalized': P,
        'cam_poses_rel_cv2': rel, 'cam_poses_rel_every2_cv2': every2,
        'seq_name': ['synthetic_%d' % b for b in range(n_objects)],
    }


def shard_objects(n_objects, rank, world_size):
    """Contiguous shard [lo, hi) of independent objects for one rank (objects never interact)."""
    base, rem = divmod(n_objects, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
