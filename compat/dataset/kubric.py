"""Synthetic stand-in for the reference's Kubric dataset class (reference dataset/kubric.py): the constructor signature, the
per-item dictionary (reference :390-402) and the canonical-frame helpers (:448-452) the scripts rely on, filled with seeded
random images on ring cameras.  Number of sequences: FORGE_SYNTHETIC_SEQS (default 16 for training, 4 otherwise)."""
import os

import torch
from torch.utils.data import Dataset

from forge_b200 import synthetic as syn


class Kubric(Dataset):
    def __init__(self, config, split='train'):
        self.config, self.split = config, split
        side = int(config.dataset.img_size)
        self.image_height, self.image_width = side, side
        every_frame = (config.test.compute_metric and split != 'train') or config.dataset.train_all_frame
        self.num_frames_per_seq = 10 if every_frame else int(config.dataset.num_frame)
        # view 0 sits on the optical axis at distance camera_z (reference dataset/kubric.py:100-103)
        ext = torch.eye(4)
        ext[2, 3] = float(config.render.camera_z)
        self._canonical_ext = ext
        self._canonical_pose = torch.linalg.inv(ext)
        default = 16 if split == 'train' else 4
        count = int(os.environ.get("FORGE_SYNTHETIC_SEQS", default))
        self._names = ["synthetic_{}_{:04d}".format(split, i) for i in range(count)]

    def __len__(self):
        return len(self._names)

    def __getitem__(self, idx):
        seed = idx if self.split == 'train' else idx + 100000
        batch = syn.kubric_batch(1, n_views_all=self.num_frames_per_seq, img_size=self.image_height,
                                 camera_z=self.config.render.camera_z, seed=seed)
        item = {key: value[0] for key, value in batch.items()}
        frames = self.num_frames_per_seq
        item['depths'] = torch.zeros(frames, 1, self.image_height, self.image_width)
        item['images'] = item['images'] * item['fg_probabilities']          # black background outside the mask
        item['seq_name'] = self._names[idx]
        if self.split == 'test':
            item['seen_flag'] = 1
        return item

    def get_canonical_extrinsics_cv2(self, device='cpu'):
        return self._canonical_ext.to(device)

    def get_canonical_pose_cv2(self, device='cpu'):
        return self._canonical_pose.to(device)

    # attribute names some reference code paths read directly
    @property
    def canonical_extrinsics_cv2(self):
        return self._canonical_ext

    @property
    def canonical_pose_cv2(self):
        return self._canonical_pose
