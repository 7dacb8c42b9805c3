"""Volumes (restatement of pytorch3d.structures.Volumes, 0.7.0, for cubic/rectangular grids).

Conventions (published PyTorch3D behaviour):
  * local coordinates span [-1, 1] over voxel CENTRES of the first/last voxel on each axis
    (i.e. ``align_corners=True`` for ``grid_sample``), last-dim order (x, y, z) <-> (W, H, D);
  * local -> world is ``Scale((size_xyz - 1) * voxel_size / 2)`` then ``Translate(-translation)``.
"""
import torch
from ..transforms import Transform3d


class Volumes:
    def __init__(self, densities, features=None, voxel_size=1.0, volume_translation=(0.0, 0.0, 0.0)):
        if densities.dim() != 5:
            raise ValueError("densities must be [N, 1|C, D, H, W]")
        if features is not None and features.shape[0] != densities.shape[0]:
            raise ValueError("features and densities need the same batch size")
        self._densities = densities
        self._features = features
        self.device = densities.device
        n = densities.shape[0]
        vs = torch.as_tensor(voxel_size, dtype=torch.float32, device=self.device)
        if vs.dim() == 0:
            vs = vs.reshape(1, 1).expand(n, 3)
        elif vs.dim() == 1:
            vs = vs.reshape(1, -1).expand(n, 3)
        self._voxel_size = vs.clone()
        tr = torch.as_tensor(volume_translation, dtype=torch.float32, device=self.device)
        self._translation = tr.reshape(1, 3).expand(n, 3).clone() if tr.dim() == 1 else tr

    def densities(self):
        return self._densities

    def features(self):
        return self._features

    def get_grid_sizes(self):
        d, h, w = self._densities.shape[2:]
        return torch.tensor([[d, h, w]], dtype=torch.int64, device=self.device).expand(self._densities.shape[0], 3)

    def get_local_to_world_coords_transform(self):
        size_xyz = self.get_grid_sizes().float()[:, [2, 1, 0]]
        scale = (size_xyz - 1) * self._voxel_size * 0.5
        return Transform3d.scale(scale).compose(Transform3d.translate(-self._translation))

    def get_world_to_local_coords_transform(self):
        return self.get_local_to_world_coords_transform().inverse()

    def world_to_local_coords(self, pts_world):
        shp = pts_world.shape
        out = self.get_world_to_local_coords_transform().transform_points(pts_world.reshape(shp[0], -1, 3))
        return out.reshape(shp)

    def local_to_world_coords(self, pts_local):
        shp = pts_local.shape
        out = self.get_local_to_world_coords_transform().transform_points(pts_local.reshape(shp[0], -1, 3))
        return out.reshape(shp)

    def get_coord_grid(self, world_coordinates=True):
        n, _, d, h, w = self._densities.shape
        axes = [torch.linspace(-1.0, 1.0, r, dtype=torch.float32, device=self.device) for r in (d, h, w)]
        Z, Y, X = torch.meshgrid(*axes, indexing="ij")
        grid = torch.stack((X, Y, Z), dim=3)[None].repeat(n, 1, 1, 1, 1)   # [N, D, H, W, 3] in (x, y, z)
        if world_coordinates:
            grid = self.local_to_world_coords(grid)
        return grid
