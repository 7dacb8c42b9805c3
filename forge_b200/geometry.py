"""Small camera helpers the reference scripts take from PyTorch3D (host math, no kernels)."""
import math

import torch
import torch.nn.functional as F


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0, degrees=True, eye=None, at=((0, 0, 0),), up=((0, 1, 0),),
                           device="cpu"):
    """Cameras on a sphere of radius ``dist`` around ``at`` (elevation / azimuth in degrees by default), looking at it.
    Returns (R [N,3,3], T [N,3]) in PyTorch3D's row-vector convention X_view = X_world @ R + T -- the call the
    reference uses for its 360-degree renders (demo.py:87, kubric_eval.py:194, scripts/kubric_validation.py:43)."""
    at_t = torch.as_tensor(at, dtype=torch.float32, device=device).reshape(-1, 3)
    up_t = torch.as_tensor(up, dtype=torch.float32, device=device).reshape(-1, 3)
    if eye is not None:
        C = torch.as_tensor(eye, dtype=torch.float32, device=device).reshape(-1, 3)
    else:
        d, e, a = [torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1) for v in (dist, elev, azim)]
        n = max(d.numel(), e.numel(), a.numel())
        d, e, a = d.expand(n), e.expand(n), a.expand(n)
        if degrees:
            e, a = e * (math.pi / 180.0), a * (math.pi / 180.0)
        C = torch.stack([d * torch.cos(e) * torch.sin(a), d * torch.sin(e), d * torch.cos(e) * torch.cos(a)], dim=1) + at_t
    z_axis = F.normalize(at_t - C, eps=1e-5)
    x_axis = F.normalize(torch.cross(up_t.expand_as(z_axis), z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    degenerate = torch.isclose(x_axis, torch.zeros(()), atol=5e-3).all(dim=1, keepdim=True)
    if degenerate.any():          # looking straight along `up`
        x_axis = torch.where(degenerate, F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5), x_axis)
    R = torch.stack([x_axis, y_axis, z_axis], dim=1).transpose(1, 2)
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R, T
