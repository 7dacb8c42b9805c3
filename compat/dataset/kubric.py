"""Synthetic stand-in for reference dataset/kubric.py (class Kubric): same constructor, same per-item dict
(dataset/kubric.py:390-402), same canonical-frame helpers (:448-452), seeded random images on ring cameras.
Sequence count via FORGE_SYNTHETIC_SEQS (default 16 train / 4 test)."""
import os

import torch
from torch.utils.data import Dataset

from forge_b200 import synthetic as syn


class Kubric(Dataset):
    def __init__(self, config, split='train'):
        self.config = config
        self.split = split
        self.image_height = self.image_width = config.dataset.img_size
        self.num_frames_per_seq = 10 if ((config.test.compute_metric and split != 'train') or config.dataset.train_all_frame) \
            else config.dataset.num_frame
        self.canonical_extrinsics_cv2 = torch.tensor([[1.0, 0.0, 0.0, 0.0],
                                                      [0.0, 1.0, 0.0, 0.0],
                                                      [0.0, 0.0, 1.0, config.render.camera_z],
                                                      [0.0, 0.0, 0.0, 1.0]])
        self.canonical_pose_cv2 = torch.inverse(self.canonical_extrinsics_cv2)
        n = int(os.environ.get("FORGE_SYNTHETIC_SEQS", "16" if split == 'train' else "4"))
        self.seq_names = ['synthetic_%s_%04d' % (split, i) for i in range(n)]

    def __len__(self):
        return len(self.seq_names)

    def __getitem__(self, idx):
        b = syn.kubric_batch(1, n_views_all=self.num_frames_per_seq, img_size=self.image_height,
                             camera_z=self.config.render.camera_z, seed=idx + (0 if self.split == 'train' else 100000))
        sample = {k: (v[0] if torch.is_tensor(v) else v[0]) for k, v in b.items()}
        sample['depths'] = torch.zeros(self.num_frames_per_seq, 1, self.image_height, self.image_width)
        sample['images'] = sample['images'] * sample['fg_probabilities']        # black background, like mask_images=True
        sample['seq_name'] = self.seq_names[idx]
        if self.split == 'test':
            sample['seen_flag'] = 1
        return sample

    def get_canonical_extrinsics_cv2(self, device='cpu'):
        return self.canonical_extrinsics_cv2.to(device)

    def get_canonical_pose_cv2(self, device='cpu'):
        return self.canonical_pose_cv2.to(device)
