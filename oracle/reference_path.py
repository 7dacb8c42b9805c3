"""TEST INFRASTRUCTURE: the reference's call sequence for the hot path, over the PyTorch3D stand-in.

Every function cites the reference lines it follows.  These issue the same ATen ops in the
same order as the reference (+ PyTorch3D 0.7.0), so run on CPU they are the "reference CPU
path" of BASELINE.json, and run on CUDA they are the "reference PyTorch3D GPU renderer".
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import enable_p3d_standin

enable_p3d_standin()
from pytorch3d.structures import Volumes  # noqa: E402
from pytorch3d.renderer import VolumeRenderer, NDCGridRaysampler, EmissionAbsorptionRaymarcher  # noqa: E402
from pytorch3d.utils.camera_conversions import cameras_from_opencv_projection  # noqa: E402


def make_conv_rgb(k_size=5):
    """models/volume_render.py:27-37 -- the 2-D decoder applied to the composited feature image."""
    pad = k_size // 2
    return nn.Sequential(
        nn.ConvTranspose2d(16, 16, kernel_size=k_size + 1, stride=2, padding=pad),
        nn.BatchNorm2d(16),
        nn.LeakyReLU(inplace=True),
        nn.Conv2d(16, 8, kernel_size=k_size, stride=1, padding=pad),
        nn.BatchNorm2d(8),
        nn.LeakyReLU(inplace=True),
        nn.Conv2d(8, 3, kernel_size=k_size, stride=1, padding=pad),
    )


def make_renderer(img_size, n_pts_per_ray, min_depth, max_depth):
    """models/volume_render.py:18-24."""
    rs = NDCGridRaysampler(image_width=img_size // 2, image_height=img_size // 2,
                           n_pts_per_ray=n_pts_per_ray, min_depth=min_depth, max_depth=max_depth)
    return VolumeRenderer(raysampler=rs, raymarcher=EmissionAbsorptionRaymarcher())


def make_cameras(camera_params, img_size, B, device):
    """models/volume_render.py:50-56 (mutates camera_params['K'] in place, as the reference does)."""
    camera_params['K'] /= 2.0
    camera_params['K'][:, -1, -1] = 1.0
    return cameras_from_opencv_projection(
        R=camera_params['R'], tvec=camera_params['T'], camera_matrix=camera_params['K'],
        image_size=torch.tensor([img_size // 2] * 2).unsqueeze(0).repeat(B, 1)).to(device)


def raymarch_core(renderer, camera_params, feature_3d, density_3d, img_size, volume_size, render_depth=False):
    """models/volume_render.py:47-63: cameras + Volumes + VolumeRenderer -> [B, S, S, C+1(+1)]."""
    B, C, D, H, W = feature_3d.shape
    cameras = make_cameras(camera_params, img_size, B, feature_3d.device)
    volume = Volumes(densities=density_3d, features=feature_3d, voxel_size=volume_size / D)
    rendered = renderer(cameras=cameras, volumes=volume, render_depth=render_depth)[0]
    return rendered, cameras


def volrender_forward(renderer, conv_rgb, camera_params, feature_3d, density_3d, img_size, volume_size,
                      render_depth=False, return_origin_proj=False):
    """models/volume_render.py:40-88 (VolRender.forward), same return-tuple ordering."""
    B, C, D, H, W = feature_3d.shape
    device = feature_3d.device
    rendered, cameras = raymarch_core(renderer, camera_params, feature_3d, density_3d, img_size, volume_size,
                                      render_depth)
    if not render_depth:
        imgs, sil = rendered.split([C, 1], dim=-1)
    else:
        imgs, sil, depth = rendered.split([C, 1, 1], dim=-1)
        depth = depth.permute(0, 3, 1, 2).contiguous()
        depth = F.interpolate(depth, size=[img_size] * 2, mode='bilinear')   # F.upsample == interpolate
    imgs = imgs.permute(0, 3, 1, 2).contiguous()
    sil = sil.permute(0, 3, 1, 2).contiguous()
    imgs = F.relu(conv_rgb(imgs))
    sil = F.interpolate(sil, size=[img_size] * 2, mode='bilinear')
    out = [imgs, sil]
    if render_depth:
        out.append(depth)
    if return_origin_proj:
        origin = torch.zeros(1, 3).to(device)
        out.append(cameras.transform_points_screen(origin, eps=1e-6).squeeze()[:, :2])
    return tuple(out)


def proj_origin(camera_params, img_size, device):
    """models/volume_render.py:91-103."""
    B = camera_params['K'].shape[0]
    cameras = make_cameras(camera_params, img_size, B, device)
    origin = torch.zeros(1, 3).to(device)
    return cameras.transform_points_screen(origin, eps=1e-6).squeeze()[:, :2]


# --------------------------------------------------------------------------------------------
def rotate_grid(grid_size, volume_size):
    """models/rotate.py:48-52 (_compute_grid): world location of each voxel centre, [D, H, W, 3]."""
    density = torch.zeros(1, 1, grid_size, grid_size, grid_size)
    volume = Volumes(densities=density, voxel_size=volume_size / grid_size)
    return volume.get_coord_grid(world_coordinates=True).squeeze()


def rotate_transformation(camPoses_cv2):
    """models/rotate.py:64-89 (get_transformation): T = pose_0 @ inverse(pose_1)."""
    B, t = camPoses_cv2.shape[:2]
    pose_0 = camPoses_cv2[:, 0:1].repeat(1, t - 1, 1, 1).reshape(B * (t - 1), 4, 4)
    pose_1 = camPoses_cv2[:, 1:].reshape(B * (t - 1), 4, 4)
    return pose_0 @ torch.inverse(pose_1)


def rotate_world_forward(voxels, camPoses_cv2, grid_size, volume_size=1.0):
    """models/rotate.py:92-156 (Rotate_world.forward); note grid_sample's default align_corners=False
    at :137 and the view-0 passthrough at :141."""
    B, t, C, D, H, W = voxels.shape
    device = voxels.device
    grid_coord = rotate_grid(grid_size, volume_size)
    grid_coord_max = grid_coord.max().item()
    T = rotate_transformation(camPoses_cv2)
    pos = grid_coord.unsqueeze(0).repeat(B * (t - 1), 1, 1, 1, 1)
    ones = torch.ones(B * (t - 1), D, H, W, 1).to(pos)
    pos = torch.cat([pos, ones], dim=-1).to(device).reshape(B * (t - 1), -1, 4)
    cam = torch.matmul(pos, T.permute(0, 2, 1))[:, :, :3]
    sample_grid = (cam / grid_coord_max).reshape(B * (t - 1), D, H, W, 3)
    out = F.grid_sample(voxels[:, 1:].reshape(B * (t - 1), C, D, H, W), sample_grid, padding_mode='zeros',
                        align_corners=False)
    out = out.reshape(B, t - 1, C, D, H, W)
    return torch.cat([voxels[:, 0:1], out], dim=1)


def sequence_from_distance(trans):
    """models/model.py:152-158."""
    dist = ((trans - trans[:, 0:1, :]) ** 2).sum(dim=-1)
    return torch.sort(dist, descending=False)[1]


def chose_selected(tensor, idxs):
    """models/model.py:161-168."""
    return torch.stack([tensor[i][idxs[i]] for i in range(len(idxs))])
