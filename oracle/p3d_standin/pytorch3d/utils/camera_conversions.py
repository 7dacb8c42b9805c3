"""OpenCV (R, tvec, K, image_size) -> PyTorch3D NDC PerspectiveCameras (0.7.0 semantics).

  scale = min(W, H) / 2;  f_ndc = (K00, K11) / scale;  p_ndc = -((K02, K12) - (W, H)/2) / scale
  R_p3d = R^T with the first two COLUMNS negated;  T_p3d = tvec with x, y negated
"""
import torch
from ..renderer.cameras import PerspectiveCameras


def cameras_from_opencv_projection(R, tvec, camera_matrix, image_size):
    focal = torch.stack([camera_matrix[:, 0, 0], camera_matrix[:, 1, 1]], dim=-1)
    principal = camera_matrix[:, :2, 2]
    size_wh = image_size.to(R).flip(dims=(1,))
    scale = (size_wh.min(dim=1, keepdim=True)[0] / 2.0).expand(-1, 2)
    c0 = size_wh / 2.0
    focal_ndc = focal / scale
    p0_ndc = -(principal - c0) / scale
    R_p3d = R.clone().permute(0, 2, 1)
    T_p3d = tvec.clone()
    R_p3d[:, :, :2] *= -1
    T_p3d[:, :2] *= -1
    return PerspectiveCameras(R=R_p3d, T=T_p3d, focal_length=focal_ndc, principal_point=p0_ndc,
                              image_size=image_size, device=R.device)
