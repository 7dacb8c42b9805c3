"""PerspectiveCameras in NDC + look_at_view_transform (restatement, PyTorch3D 0.7.0 semantics).

Row-vector conventions: ``X_view = X_world @ R + T``; projection matrix (column form)
``[[fx,0,px,0],[0,fy,py,0],[0,0,0,1],[0,0,1,0]]`` so that ndc = (fx X/Z + px, fy Y/Z + py, 1/Z).
PyTorch3D's NDC/screen axes: +x left, +y up.
"""
import math
import torch
import torch.nn.functional as F
from ..transforms import Transform3d


class PerspectiveCameras:
    def __init__(self, focal_length=1.0, principal_point=((0.0, 0.0),), R=None, T=None,
                 device="cpu", in_ndc=True, image_size=None):
        if R is None:
            R = torch.eye(3)[None]
        if T is None:
            T = torch.zeros(1, 3)
        self.R = R.to(device)
        self.T = T.to(device)
        n = self.R.shape[0]
        fl = torch.as_tensor(focal_length, dtype=torch.float32, device=device)
        if fl.dim() == 0:
            fl = fl.reshape(1, 1).expand(n, 2)
        elif fl.dim() == 1:
            fl = fl.reshape(-1, 1).expand(-1, 2)
        self.focal_length = fl
        self.principal_point = torch.as_tensor(principal_point, dtype=torch.float32, device=device)
        self.image_size = None if image_size is None else torch.as_tensor(image_size, device=device)
        self._in_ndc = in_ndc
        self._N = n
        self.device = torch.device(device) if not isinstance(device, torch.device) else device

    def in_ndc(self):
        return self._in_ndc

    def to(self, device):
        return PerspectiveCameras(focal_length=self.focal_length, principal_point=self.principal_point,
                                  R=self.R, T=self.T, device=device, in_ndc=self._in_ndc,
                                  image_size=self.image_size)

    def get_image_size(self):
        return self.image_size

    # -- transforms ------------------------------------------------------------------------
    def get_world_to_view_transform(self):
        return Transform3d.rotate(self.R).compose(Transform3d.translate(self.T))

    def get_projection_transform(self):
        n = self._N
        K = torch.zeros(n, 4, 4, dtype=torch.float32, device=self.R.device)
        K[:, 0, 0] = self.focal_length[:, 0]
        K[:, 1, 1] = self.focal_length[:, 1]
        K[:, 0, 2] = self.principal_point[:, 0]
        K[:, 1, 2] = self.principal_point[:, 1]
        K[:, 3, 2] = 1.0
        K[:, 2, 3] = 1.0
        return Transform3d(K.transpose(1, 2).contiguous())

    def get_full_projection_transform(self):
        return self.get_world_to_view_transform().compose(self.get_projection_transform())

    def unproject_points(self, xy_depth, world_coordinates=True, from_ndc=False):
        if world_coordinates:
            to_cam = self.get_full_projection_transform()
        else:
            to_cam = self.get_projection_transform()
        if not self._in_ndc and from_ndc:
            raise NotImplementedError("stand-in only models cameras defined in NDC")
        unproj = to_cam.inverse()
        xy_inv_depth = torch.cat((xy_depth[..., :2], 1.0 / xy_depth[..., 2:3]), dim=-1)
        return unproj.transform_points(xy_inv_depth)

    def transform_points_ndc(self, points, eps=None):
        return self.get_full_projection_transform().transform_points(points, eps=eps)

    def transform_points_screen(self, points, eps=None, with_xyflip=True):
        ndc = self.transform_points_ndc(points, eps=eps)
        size = self.image_size.reshape(-1, 2).to(self.R.device)
        height, width = size.unbind(1)
        n = self._N
        K = torch.zeros(n, 4, 4, dtype=torch.float32, device=self.R.device)
        s = size.min(dim=1).values.float() / 2.0
        K[:, 0, 0] = s
        K[:, 1, 1] = s
        K[:, 0, 3] = -1.0 * width.float() / 2.0
        K[:, 1, 3] = -1.0 * height.float() / 2.0
        K[:, 2, 2] = 1.0
        K[:, 3, 3] = 1.0
        to_screen = Transform3d(K.transpose(1, 2).contiguous())
        if with_xyflip:
            flip = torch.eye(4, dtype=torch.float32, device=self.R.device)
            flip[0, 0] = -1.0
            flip[1, 1] = -1.0
            to_screen = to_screen.compose(Transform3d(flip[None].expand(n, -1, -1).contiguous()))
        return to_screen.transform_points(ndc, eps=eps)


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees=True, at=((0, 0, 0),), up=((0, 1, 0),),
                           device="cpu"):
    """Camera on a sphere around ``at`` looking at it. Returns (R [N,3,3], T [N,3]) in the
    PyTorch3D row-vector convention (used by demo.py:87 / kubric_eval.py:194 for 360 renders)."""
    dist, elev, azim = [torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1) for v in (dist, elev, azim)]
    n = max(dist.numel(), elev.numel(), azim.numel())
    dist, elev, azim = [v.expand(n) for v in (dist, elev, azim)]
    if degrees:
        elev = math.pi / 180.0 * elev
        azim = math.pi / 180.0 * azim
    x = dist * torch.cos(elev) * torch.sin(azim)
    y = dist * torch.sin(elev)
    z = dist * torch.cos(elev) * torch.cos(azim)
    at_t = torch.as_tensor(at, dtype=torch.float32, device=device).reshape(-1, 3)
    up_t = torch.as_tensor(up, dtype=torch.float32, device=device).reshape(-1, 3)
    C = torch.stack([x, y, z], dim=1) + at_t
    z_axis = F.normalize(at_t - C, eps=1e-5)
    x_axis = F.normalize(torch.cross(up_t.expand_as(z_axis), z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    close = torch.isclose(x_axis, torch.tensor(0.0), atol=5e-3).all(dim=1, keepdim=True)
    if close.any():
        x_axis = torch.where(close, F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5), x_axis)
    R = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1).transpose(1, 2)
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R, T
