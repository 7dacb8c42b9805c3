"""FORGE_poseEstimator3D -- mirror of reference models/model_single_pose_estimator.py (:14-138).

Training steps 1.1 / 1.2 of the reference: renders 2t views per object from THREE distinct fused
volumes (first three views, last two views, all views; reference :104-131).  The reference
materialises 2t copies of those volumes per object; here K1 receives the 3b distinct volumes and a
``view2vol`` table.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import Encoder3D
from .volume_render import VolRender, origin_projection
from .rotate import Rotate_world
from .model import _mat2quat


class FORGE_poseEstimator3D(nn.Module):
    def __init__(self, config, pose_estimator_3d_cls=None, mat2quat=None):
        super(FORGE_poseEstimator3D, self).__init__()
        self.config = config
        self.encoder_3d = Encoder3D(config)
        self.render = VolRender(config)
        self.rotate = Rotate_world(config)
        self._mat2quat = mat2quat or _mat2quat
        if pose_estimator_3d_cls is not None:
            self.encoder_traj = pose_estimator_3d_cls(config)

    def forward(self, sample, dataset, device):
        clips = sample['images'].to(device)
        b, t, c, h, w = clips.shape
        features_raw = self.encoder_3d.get_feat3D(clips.reshape(b * t, c, h, w))
        _, C, D, H, W = features_raw.shape
        features_raw = features_raw.reshape(b, t, C, D, H, W)

        if not self.config.train.use_gt_pose:
            if not hasattr(self, 'encoder_traj'):
                raise RuntimeError("FORGE_poseEstimator3D was built without a pose estimator; pass "
                                   "pose_estimator_3d_cls or set config.train.use_gt_pose = True")
            poses_cam, conf = self.encoder_traj(features_raw)
            poses_cam = torch.cat([F.normalize(poses_cam[:, :4]), poses_cam[:, 4:]], dim=1)
            camPoseRel_cv2 = self.encoder_traj.toSE3(poses_cam)
            canonical_pose = dataset.get_canonical_pose_cv2(device=device)
            canonical_ext = dataset.get_canonical_extrinsics_cv2(device=device)
            camPoses_cv2 = canonical_pose.unsqueeze(0) @ camPoseRel_cv2
            camE_cv2 = torch.inverse(camPoses_cv2).reshape(b, t - 1, 4, 4)
            camPoses_cv2 = camPoses_cv2.reshape(b, t - 1, 4, 4)
            camPoses_cv2 = torch.cat([canonical_pose.reshape(1, 1, 4, 4).repeat(b, 1, 1, 1), camPoses_cv2], dim=1)
            camE_cv2 = torch.cat([canonical_ext.reshape(1, 1, 4, 4).repeat(b, 1, 1, 1), camE_cv2], dim=1)
            gt = self._mat2quat(sample['cam_poses_rel_cv2'][:, 1:].to(device).reshape(b * (t - 1), 4, 4))
            camPose_return = {'gt': gt, 'pred': poses_cam, 'conf': conf}
        else:
            key = '_canonicalized' if self.config.train.canonicalize else ''
            camE_cv2 = sample['cam_extrinsics_cv2' + key].to(device)
            camPoses_cv2 = sample['cam_poses_cv2' + key].to(device)
            camPose_return = None

        # every view is rendered twice: once from a partial-view volume, once from the all-view volume
        camE2 = camE_cv2.repeat(1, 2, 1, 1).reshape(b * 2 * t, 4, 4)
        cameras = {'R': camE2[:, :3, :3], 'T': camE2[:, :3, 3],
                   'K': sample['K_cv2'].repeat(1, 2, 1, 1).reshape(b * 2 * t, 3, 3).to(device)}

        if self.config.train.parameter == 'pose':
            K = cameras['K']
            K /= 2.0
            K[:, -1, -1] = 1.0
            origin_proj = origin_projection(cameras['T'], K)
            # this branch of the reference keeps the third (1/z) screen coordinate (reference :97-98)
            origin_proj = torch.cat([origin_proj, 1.0 / cameras['T'][:, 2:3]], dim=1)
            return camPose_return, 2 * origin_proj / self.config.dataset.img_size

        features_transformed = self.rotate(voxels=features_raw, camPoses_cv2=camPoses_cv2[:, :t], grid_size=D)
        f_3v = self.encoder_3d.fuse(features_transformed[:, :3])           # 3-view volume
        f_2v = self.encoder_3d.fuse(features_transformed[:, -2:])          # 2-view volume
        f_mv = self.encoder_3d.fuse(features_transformed)                  # all-view volume
        if self.encoder_3d.density_head.training or self.encoder_3d.features_head.training:
            # batch-statistics BN: the reference runs the heads on the 2b partial-view volumes and on the b all-view
            # volumes in separate calls (reference :110-111, :123-124) -- same statistics, same running-stat updates
            part = torch.cat([f_3v, f_2v], dim=0)
            densities = torch.cat([self.encoder_3d.get_density3D(part), self.encoder_3d.get_density3D(f_mv)], dim=0)
            features = torch.cat([self.encoder_3d.get_render_features(part),
                                  self.encoder_3d.get_render_features(f_mv)], dim=0)
        else:                                                              # eval-mode BN is per-sample: one 3b call
            fused = torch.cat([f_3v, f_2v, f_mv], dim=0)
            densities = self.encoder_3d.get_density3D(fused)
            features = self.encoder_3d.get_render_features(fused)
        if self.config.dataset.name == 'omniobject3d':
            densities = densities.clamp(min=0.0, max=1.0)

        # reference view order per object: [2v-volume x 3 views, 3v-volume x 2 views, all-view volume x t views]
        assert t == 5, "the 3-view / 2-view split of the reference is defined for 5 input views"
        per_obj = torch.tensor([1] * 3 + [0] * 2 + [2] * t, dtype=torch.int64)
        view2vol = (per_obj.view(1, -1) * b + torch.arange(b).view(b, 1)).reshape(-1).int().to(device)
        rendered_imgs, rendered_masks, origin_proj = self.render(cameras, features, densities,
                                                                 return_origin_proj=True, view2vol=view2vol)
        if self.config.train.use_gt_pose:
            return rendered_imgs, rendered_masks
        return rendered_imgs, rendered_masks, 2 * origin_proj / self.config.dataset.img_size, camPose_return
