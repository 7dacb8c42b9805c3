"""The reference's OWN demo.py and kubric_train_pose_3D.py, unmodified, executed end to end against forge_b200 through compat/ (BASELINE.json north_star:
"kubric_train_*.py and demo.py run unchanged").  Needs the staged reference tarball (tools/stage_reference.sh ->
baseline/_ref/forge_reference.tar.gz, git-ignored) or a checkout at $FORGE_REFERENCE; skipped otherwise.  The script predicts
poses, runs its 3 x 2001 refinement iterations through rotate -> fuse -> heads -> render -> backward, renders 28 novel views
per case and writes GIFs; weights are random-init synthetic checkpoints with the reference's state_dict keys."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "forge_reference.tar.gz")) or \
    os.path.isdir(os.path.join(os.environ.get("FORGE_REFERENCE", "/root/reference"), "models"))


@pytest.mark.skipif(not HAVE, reason="no reference checkout / staged tarball")
def test_reference_demo_py_runs_unchanged():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "demo"],
                         capture_output=True, text=True, timeout=1500)
    last = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    summary = json.loads(last[-1])
    assert summary["rc"] == 0
    assert summary["gifs"] == ["0_0.gif", "1_0.gif", "2_0.gif"]          # one 360-degree render per demo case


@pytest.mark.skipif(not HAVE, reason="no reference checkout / staged tarball")
def test_reference_kubric_train_pose_3d_runs_unchanged():
    """The reference's kubric_train_pose_3D.py (training step 1.1: config/kubric/gt_pose.yaml, FORGE_poseEstimator3D with every
    parameter trainable, VGG perceptual loss, SyncBatchNorm + DDP) launched with torchrun on one GPU, unmodified apart from the
    run length of its yaml: 2 x 4 iterations through lift -> rotate -> ConvGRU -> heads -> render (30 views per object) ->
    backward (K1 / K2 gradients to the volumes) -> Adam, loss printed finite every iteration."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "train_pose3d", "--gpus", "1",
                          "--iters", "4", "--batch", "4"], capture_output=True, text=True, timeout=900)
    last = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    summary = json.loads(last[-1])
    assert summary["rc"] == 0
    losses = [float(l.split("recon_img_mv: ")[1].split(" ")[0]) for l in summary["log_lines"] if "recon_img_mv" in l]
    assert len(losses) >= 4 and all(0.0 < x < 10.0 for x in losses)


JOINT = r'''
import sys, warnings, numpy as np, torch
warnings.simplefilter("ignore")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from types import SimpleNamespace as NS
from oracle import seeded
from forge_b200 import synthetic as syn
from models.model import FORGE                      # compat: forge_b200 FORGE + the reference's own pose networks
seed = int(sys.argv[1])
cfg = syn.make_config(img_size=256, n_pts_per_ray=32, use_gt_pose=False, parameter='joint')
cfg.network.rot_representation = 'quat'
m = seeded.load_seeded(FORGE(cfg), seed).cuda().eval()
m.encoder_3d.density_head[6].bias.data.fill_(0.1)
sample = syn.kubric_batch(1, n_views_all=7, img_size=256, seed=seed)
class DS:
    ext = torch.eye(4); ext[2, 3] = cfg.render.camera_z
    def get_canonical_extrinsics_cv2(self, device='cpu'): return self.ext.to(device)
    def get_canonical_pose_cv2(self, device='cpu'): return torch.inverse(self.ext).to(device)
with torch.no_grad():
    rgb, mask, oproj, poses = m(sample, DS(), 'cuda')
np.savez(sys.argv[2], rgb_sub=seeded.subsample(rgb).cpu().numpy(), mask_sub=seeded.subsample(mask).cpu().numpy(),
         origin_proj=oproj.cpu().numpy(), pose_pred=poses['pred'].cpu().numpy(), pose_gt=poses['gt'].cpu().numpy(),
         pose_conf=poses['conf'].cpu().numpy())
'''


@pytest.mark.skipif(not HAVE, reason="no reference checkout / staged tarball (the pose networks are the reference's own)")
def test_joint_model_with_predicted_poses_matches_reference_values(tmp_path):
    """FORGE.forward with PREDICTED poses (both pose networks + pose head -> canonical pose algebra -> rotate -> sorted fuse ->
    heads -> 5 input + 2 novel views) against the output of the unmodified reference FORGE on CPU (tests/golden/
    joint_model_eval.npz, oracle/make_golden.py); seeded weights on both sides."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_reference_script as rrs
    ref = rrs.stage(str(tmp_path))
    g = np.load(os.path.join(ROOT, "tests", "golden", "joint_model_eval.npz"))
    out = str(tmp_path / "joint.npz")
    res = subprocess.run([sys.executable, "-c", JOINT, str(int(g['seed'])), out], cwd=ref, env=rrs.env_for(ref),
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    o = np.load(out)
    # the pose networks are the reference's code on both sides: predictions agree to fp32 (GPU vs CPU) accuracy
    assert np.abs(o['pose_pred'] - g['pose_pred']).max() <= 2e-3
    assert np.abs(o['pose_gt'] - g['pose_gt']).max() <= 1e-5
    assert np.abs(o['origin_proj'] - g['origin_proj']).max() <= 5e-3
    for k in ('rgb_sub', 'mask_sub'):
        ref_v = g[k]
        assert o[k].shape == ref_v.shape
        assert np.abs(o[k] - ref_v).max() <= 2e-2 * max(1.0, np.abs(ref_v).max()), k       # images through predicted poses
        assert np.abs(o[k] - ref_v).mean() <= 1e-3, k
    assert g['mask_sub'].max() > 0.5
