"""Rotate_world -- drop-in for reference models/rotate.py (class Rotate_world, :9-156).

Same constructor, ``forward(voxels, camPoses_cv2, grid_size)`` signature, parameters
(``conv3d_1..4``, unused by the live path but present in checkpoints) and plain-attribute grid
tables.  The materialised grid + matmul + ``F.grid_sample`` + ``torch.cat`` (reference :127-141)
are one CUDA launch (``forge_rotate_fwd/bwd``); ``T = pose_0 @ inverse(pose_1)`` is one more small
launch (``forge_pose_affine_fwd``, differentiable so gradients reach the camera poses).
"""
import torch
import torch.nn as nn

from .. import ops


def normal_init(module, mean=0, std=1, bias=0):
    """reference utils/train_utils.py:143-146"""
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias'):
        nn.init.constant_(module.bias, bias)


class Rotate_world(nn.Module):
    '''
    Rotate the voxel in world coordinate using camera relative poses
    '''
    def __init__(self, config):
        super(Rotate_world, self).__init__()
        self.padding_mode = config.network.padding_mode   # stored, ignored: the resample is 'zeros' (ref :138)

        self.grid_size = 32
        self.vol_size = config.render.volume_size
        self.single_voxel_size = self.vol_size / self.grid_size

        # world location of voxel centres along one axis (the reference keeps the full [D,H,W,3] grid;
        # it is separable) and its max = volume half size, 0.4844 for 32 (ref :22-35)
        self._axis = {}
        self._dev_axes = {}
        self._job_cache = {}
        self.grid_coord_max = self._compute_axis(32)[1]
        self.grid_coord_max_16 = self._compute_axis(16)[1]
        self.grid_coord_max_64 = self._compute_axis(64)[1]
        self.grid_coord_max_128 = self._compute_axis(128)[1]
        self.grid_coord_max_48 = self._compute_axis(48)[1]

        self.conv3d_1 = nn.Conv3d(16, 16, 3, padding=1)
        self.conv3d_2 = nn.Conv3d(16, 16, 3, padding=1)
        normal_init(self.conv3d_1, mean=0.0, std=0.01, bias=0)
        normal_init(self.conv3d_2, mean=0.0, std=0.01, bias=0)

        self.conv3d_3 = nn.Conv3d(128, 128, 3, padding=1)
        self.conv3d_4 = nn.Conv3d(128, 128, 3, padding=1)
        normal_init(self.conv3d_3, mean=0.0, std=0.01, bias=0)
        normal_init(self.conv3d_4, mean=0.0, std=0.01, bias=0)

    def _compute_axis(self, n):
        """fp32 voxel-centre coordinates of an n-voxel axis, rounded like PyTorch3D's
        Volumes.get_coord_grid (linspace(-1,1,n) times (n-1)*voxel_size*0.5), and their max."""
        if n not in self._axis:
            voxel = torch.tensor(self.vol_size / n, dtype=torch.float32)
            scale = (torch.tensor(float(n), dtype=torch.float32) - 1) * voxel * 0.5
            g = torch.linspace(-1.0, 1.0, n, dtype=torch.float32) * scale
            self._axis[n] = (g, g.max().item())
        return self._axis[n]

    def _device_axes(self, D, H, W, device):
        key = (D, H, W, str(device))
        if key not in self._dev_axes:
            gx, gmax = self._compute_axis(W)
            self._dev_axes[key] = (gx.to(device), self._compute_axis(H)[0].to(device),
                                   self._compute_axis(D)[0].to(device), gmax)
        return self._dev_axes[key]

    def _jobs(self, B, t, device, order):
        """(src, dst, kind) per view: view 0 of each object is a copy (reference :141), the rest resample."""
        key = (B, t, str(device))
        if key not in self._job_cache:
            src = torch.arange(B * t, dtype=torch.int32)
            kind = (src % t == 0).int()
            self._job_cache[key] = (src.to(device), kind.to(device), torch.stack([src, src, kind], dim=1).to(device))
        src, kind, plain = self._job_cache[key]
        if order is None:
            return plain
        inv = torch.argsort(order.to(device), dim=1)                         # view v lands in slot inv[b,v]
        dst = (inv + torch.arange(B, device=device).view(B, 1) * t).reshape(-1).int()
        return torch.stack([src, dst, kind], dim=1).contiguous()

    @property
    def grid_coord(self):
        """[D,H,W,3] grid like the reference attribute (built on demand; the kernel does not need it)."""
        g = self._compute_axis(self.grid_size)[0]
        Z, Y, X = torch.meshgrid(g, g, g, indexing='ij')
        return torch.stack((X, Y, Z), dim=-1)

    def get_transformation(self, camPoses_cv2):
        '''
        camPoses_cv2: [B,t,4,4].  T = pose_0 @ inverse(pose_1) (reference :64-89)
        '''
        B, t, _, _ = camPoses_cv2.shape
        pose_0 = camPoses_cv2[:, 0:1].repeat(1, t - 1, 1, 1).reshape(B * (t - 1), 4, 4)
        pose_1 = camPoses_cv2[:, 1:].reshape(B * (t - 1), 4, 4)
        return pose_0 @ torch.inverse(pose_1)

    def _jobs_aliased(self, B, t, device, order):
        """resample jobs only (no view-0 copy): view v >= 1 of object b lands in slot (position of v in order[b]) - 1 of a
        [B, t-1] output; order[b, 0] must be 0"""
        src = (torch.arange(B, device=device).view(B, 1) * t + torch.arange(1, t, device=device).view(1, t - 1))
        if order is None:
            slot = torch.arange(t - 1, device=device).view(1, t - 1).expand(B, t - 1)
        else:
            slot = torch.argsort(order.to(device), dim=1)[:, 1:] - 1
        dst = slot + torch.arange(B, device=device).view(B, 1) * (t - 1)
        return torch.stack([src.reshape(-1), dst.reshape(-1), torch.zeros(B * (t - 1), dtype=torch.long, device=device)],
                           dim=1).int().contiguous()

    def forward_views(self, voxels, camPoses_cv2, grid_size=32, order=None):
        """Same resample as ``forward`` but WITHOUT the view-0 passthrough copy (reference :141 materialises it with
        torch.cat): returns a list of t tensors [B,C,D,H,W] (channels-last memory) -- slot 0 is view 0 of the input itself
        (zero-copy when the input is channels-last), slots 1.. are views of one [B, t-1, D, H, W, C] resample output.
        ``order`` as in forward; its first column must be 0 (view 0 is the nearest view to itself)."""
        if not voxels.is_cuda:
            raise RuntimeError("forge_b200.Rotate_world needs CUDA voxels; there is no CPU path")
        B, t, C, D, H, W = voxels.shape
        device = voxels.device
        gx, gy, gz, gmax = self._device_axes(D, H, W, device)
        if grid_size != D:
            gmax = self._compute_axis(grid_size)[1]
        A = ops.pose_affine(camPoses_cv2.to(device))
        vox_cl = ops.to_channels_last(voxels.reshape(B * t, C, D, H, W))
        views = [vox_cl.view(B, t, D, H, W, C)[:, 0].permute(0, 4, 1, 2, 3)]
        if t > 1:
            A = A.view(B, t, 12)[:, 1:].reshape(B * (t - 1), 12)          # one affine per JOB (the kernel's contract)
            out_cl = ops.rotate_resample(vox_cl, A, self._jobs_aliased(B, t, device, order), gx, gy, gz, gmax, B * (t - 1))
            out = out_cl.view(B, t - 1, D, H, W, C).permute(0, 1, 5, 2, 3, 4)
            views += [out[:, i] for i in range(t - 1)]
        return views

    def forward(self, voxels, camPoses_cv2, grid_size=32, order=None):
        '''
        voxels: [B,t,C,D,H,W] features of all frames; camPoses_cv2: [B,t,4,4] camera-to-world poses.
        Returns [B,t,C,D,H,W]: view 0 untouched, views 1.. resampled into view 0's frame.  The result
        is channels-last in memory (a permuted view), which the 3-D convs downstream consume directly.
        order (optional, int [B,t]): output slot p of object b receives view order[b,p] -- the
        distance-sorted permutation of reference models/model.py:152-168 folded into the same launch.
        '''
        if not voxels.is_cuda:
            raise RuntimeError("forge_b200.Rotate_world needs CUDA voxels; there is no CPU path")
        B, t, C, D, H, W = voxels.shape
        device = voxels.device
        gx, gy, gz, gmax = self._device_axes(D, H, W, device)
        if grid_size != D:
            gmax = self._compute_axis(grid_size)[1]

        A = ops.pose_affine(camPoses_cv2.to(device))                         # [B*t,12]; view 0: identity (copy job)
        jobs = self._jobs(B, t, device, order)

        vox_cl = ops.to_channels_last(voxels.reshape(B * t, C, D, H, W))
        out_cl = ops.rotate_resample(vox_cl, A, jobs, gx, gy, gz, gmax, B * t)
        return out_cl.view(B, t, D, H, W, C).permute(0, 1, 5, 2, 3, 4)
