"""Implicit volume renderer (restatement of the PyTorch3D 0.7.0 op sequence FORGE relies on).

  NDCGridRaysampler              pixel-centre NDC grid, two-plane un-projection, linspace depths
  VolumeSampler (internal)       world->local, points = o + z*d, 2x F.grid_sample(align_corners=True)
  EmissionAbsorptionRaymarcher   w_k = s_k * prod_{j<k}(1+1e-10 - s_j); sum w f; 1 - prod(1 - s)
                                 + the depth patch the reference's README.md:26-33 adds
  VolumeRenderer                 glue, returns (images [N,H,W,C+1(+1)], ray_bundle)
"""
import warnings
from collections import namedtuple

import torch
import torch.nn.functional as F

from ..transforms import Transform3d
from .cameras import PerspectiveCameras, look_at_view_transform  # noqa: F401

RayBundle = namedtuple("RayBundle", ["origins", "directions", "lengths", "xys"])


class NDCGridRaysampler(torch.nn.Module):
    def __init__(self, image_width, image_height, n_pts_per_ray, min_depth, max_depth):
        super().__init__()
        if image_width >= image_height:
            range_x, range_y = image_width / image_height, 1.0
        else:
            range_x, range_y = 1.0, image_height / image_width
        half_w = range_x / image_width
        half_h = range_y / image_height
        ys = torch.linspace(range_y - half_h, -range_y + half_h, image_height, dtype=torch.float32)
        xs = torch.linspace(range_x - half_w, -range_x + half_w, image_width, dtype=torch.float32)
        Y, X = torch.meshgrid(ys, xs, indexing="ij")
        self.register_buffer("_xy_grid", torch.stack((X, Y), dim=-1), persistent=False)
        self._n_pts_per_ray = n_pts_per_ray
        self._min_depth = min_depth
        self._max_depth = max_depth

    def forward(self, cameras, **kwargs):
        n = cameras.R.shape[0]
        device = cameras.R.device
        xy = self._xy_grid.to(device).expand(n, -1, -1, -1)
        spatial = xy.shape[1:-1]
        n_rays = spatial[0] * spatial[1]
        depths = torch.linspace(self._min_depth, self._max_depth, self._n_pts_per_ray, dtype=xy.dtype, device=device)
        zs = depths[None, None].expand(n, n_rays, self._n_pts_per_ray)
        xy_flat = xy.reshape(n, 1, n_rays, 2).expand(n, 2, n_rays, 2).reshape(n, 2 * n_rays, 2)
        plane_depth = torch.cat((xy.new_ones(n, n_rays, 1), 2.0 * xy.new_ones(n, n_rays, 1)), dim=1)
        unproj = cameras.unproject_points(torch.cat((xy_flat, plane_depth), dim=-1), from_ndc=True)
        p1 = unproj[:, :n_rays]
        p2 = unproj[:, n_rays:]
        dirs = p2 - p1
        origins = p1 - dirs
        return RayBundle(origins.reshape(n, *spatial, 3), dirs.reshape(n, *spatial, 3),
                         zs.reshape(n, *spatial, self._n_pts_per_ray), xy)


class VolumeSampler(torch.nn.Module):
    def __init__(self, volumes, sample_mode="bilinear"):
        super().__init__()
        self._volumes = volumes
        self._sample_mode = sample_mode

    def forward(self, ray_bundle, **kwargs):
        o_w, d_w, lengths = ray_bundle.origins, ray_bundle.directions, ray_bundle.lengths
        vols = self._volumes
        if vols.densities().shape[0] != o_w.shape[0]:
            raise ValueError("Input volumes have to have the same batch size as rays.")
        o_l = vols.world_to_local_coords(o_w)
        w2l = vols.get_world_to_local_coords_transform().get_matrix()
        dmat = torch.eye(4, dtype=w2l.dtype, device=w2l.device)[None].repeat(w2l.shape[0], 1, 1)
        dmat[:, :3, :3] = w2l[:, :3, :3]
        d_l = Transform3d(dmat).transform_points(d_w.reshape(lengths.shape[0], -1, 3)).reshape(d_w.shape)
        pts = o_l[..., None, :] + lengths[..., :, None] * d_l[..., None, :]
        flat = pts.reshape(pts.shape[0], -1, 1, 1, 3)
        dens = vols.densities()
        rd = F.grid_sample(dens, flat, align_corners=True, mode=self._sample_mode, padding_mode="zeros")
        rd = rd.permute(0, 2, 3, 4, 1).reshape(*pts.shape[:-1], dens.shape[1])
        feats = vols.features()
        rf = None
        if feats is not None:
            rf = F.grid_sample(feats, flat, align_corners=True, mode=self._sample_mode, padding_mode="zeros")
            rf = rf.permute(0, 2, 3, 4, 1).reshape(*pts.shape[:-1], feats.shape[1])
        return rd, rf


class EmissionAbsorptionRaymarcher(torch.nn.Module):
    def __init__(self, surface_thickness=1):
        super().__init__()
        self.surface_thickness = surface_thickness

    def forward(self, rays_densities, rays_features, eps=1e-10, **kwargs):
        with torch.no_grad():
            if rays_densities.max() > 1.0 or rays_densities.min() < 0.0:
                warnings.warn("One or more elements of rays_densities are outside of valid range (0.0, 1.0)")
        sigma = rays_densities[..., 0]
        cp = torch.cumprod((1.0 + eps) - sigma, dim=-1)
        s = self.surface_thickness
        absorption = torch.cat([torch.ones_like(cp[..., :s]), cp[..., :-s]], dim=-1)
        weights = sigma * absorption
        features = (weights[..., None] * rays_features).sum(dim=-2)
        opacities = 1.0 - torch.prod(1.0 - sigma, dim=-1, keepdim=True)
        # reference README.md:26-33 (manual patch applied to PyTorch3D's raymarching.py)
        if kwargs.get("render_depth", False) is True and "ray_bundle" in kwargs:
            lengths = kwargs["ray_bundle"].lengths[..., None]
            depths = (weights[..., None] * lengths).sum(dim=-2)
            return torch.cat((features, opacities, depths), dim=-1)
        return torch.cat((features, opacities), dim=-1)


class VolumeRenderer(torch.nn.Module):
    def __init__(self, raysampler, raymarcher, sample_mode="bilinear"):
        super().__init__()
        self.raysampler = raysampler
        self.raymarcher = raymarcher
        self._sample_mode = sample_mode

    def forward(self, cameras, volumes, **kwargs):
        sampler = VolumeSampler(volumes, sample_mode=self._sample_mode)
        ray_bundle = self.raysampler(cameras=cameras, volumetric_function=sampler, **kwargs)
        rd, rf = sampler(ray_bundle=ray_bundle, cameras=cameras, **kwargs)
        images = self.raymarcher(rays_densities=rd, rays_features=rf, ray_bundle=ray_bundle, **kwargs)
        return images, ray_bundle
