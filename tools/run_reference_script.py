#!/usr/bin/env python
"""Run one of the reference's OWN scripts, unmodified, against forge_b200 through compat/.

    python tools/run_reference_script.py demo                 # demo.py --cfg config/demo/demo.yaml
    python tools/run_reference_script.py train --gpus 2       # kubric_train_joint.py (step 3.3 yaml, SyncBN + DDP), a few iterations
    python tools/run_reference_script.py train_pose3d --gpus 2   # kubric_train_pose_3D.py (step 1.1 gt_pose.yaml: FORGE_poseEstimator3D,
                                                                 # every parameter trainable: K1 / K2 backward to the volumes)

The reference checkout comes from $FORGE_REFERENCE, /root/reference, or the staged tarball baseline/_ref/forge_reference.tar.gz
(tools/stage_reference.sh); it is copied / extracted into a scratch directory because the scripts write ./output and ./log next
to themselves.  What this script adds around the unmodified entry point: PYTHONPATH (compat/ first), synthetic checkpoints at
the paths the scripts hard-code (random-init weights with the reference's state_dict keys), and for `train` a YAML that
differs from config/kubric/joint_pose_2d3d.yaml only in run length (total_iteration, resume) -- the training script has no
flag for that.
"""
import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARBALL = os.path.join(ROOT, "baseline", "_ref", "forge_reference.tar.gz")


def stage(workdir):
    """unmodified reference tree -> workdir/FORGE"""
    dst = os.path.join(workdir, "FORGE")
    src = os.environ.get("FORGE_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/models") else None)
    if src and os.path.isdir(os.path.join(src, "models")):
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    elif os.path.exists(TARBALL):
        os.makedirs(dst)
        with tarfile.open(TARBALL) as tf:
            tf.extractall(dst)
    else:
        raise SystemExit("no reference checkout: set FORGE_REFERENCE or run tools/stage_reference.sh")
    return dst


def env_for(ref):
    env = dict(os.environ)
    env["FORGE_REFERENCE"] = ref
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, ref])
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    return env


MAKE_CKPT = r'''
import os, sys, torch, warnings
warnings.simplefilter("ignore")
from config.config import config, update_config
update_config(sys.argv[1])
from models.model import FORGE
torch.manual_seed(0)
m = FORGE(config)
m.encoder_3d.density_head[6].bias.data.fill_(0.1)      # random-init heads end in ReLU: keep the density volume non-empty
sd = m.state_dict()
def save(obj, root, name):
    os.makedirs(root, exist_ok=True)
    torch.save(obj, os.path.join(root, name))
save({'state_dict': sd}, './output/kubric/joint_pose_2d3d/pred_pose_2d3d_joint', 'cpt_best_psnr_26.340881009038913_7.545314707482719.pth.tar')
save({'state_dict': sd}, './output/kubric/gt_pose/gt_pose', 'cpt_best_psnr_31.842686198427398.pth.tar')
save({'state_dict': sd}, './output/kubric/pred_pose_2d3d/pred_pose_2d3d', 'cpt_last_114k.pth.tar')
if len(sys.argv) > 2:      # a cpt_last for `resume: True` (epoch 1: the epoch-0 validation pass is not part of the measurement)
    param = list(m.encoder_traj.parameters()) + list(m.pose_head.parameters()) + list(m.encoder_3d.fusion_feature.parameters()) + \
            list(m.encoder_3d.density_head.parameters()) + list(m.render.parameters())
    opt = torch.optim.Adam(param, lr=config.train.lr)
    save({'epoch': 1, 'state_dict': sd, 'optimizer': opt.state_dict()}, sys.argv[2], 'cpt_last.pth.tar')
print("checkpoints written")
'''


MAKE_CKPT_POSE3D = r'''
import os, sys, torch, warnings
warnings.simplefilter("ignore")
from config.config import config, update_config
update_config(sys.argv[1])
from models.model_single_pose_estimator import FORGE_poseEstimator3D
torch.manual_seed(0)
m = FORGE_poseEstimator3D(config)
m.encoder_3d.density_head[6].bias.data.fill_(0.1)      # random-init heads end in ReLU: keep the density volume non-empty
opt = torch.optim.Adam(m.parameters(), lr=config.train.lr)
os.makedirs(sys.argv[2], exist_ok=True)
torch.save({'epoch': 1, 'state_dict': m.state_dict(), 'optimizer': opt.state_dict()}, os.path.join(sys.argv[2], 'cpt_last.pth.tar'))
print("checkpoint written")
'''

TRAIN_MODES = {      # what -> (script, yaml under config/kubric, checkpoint maker)
    "train": ("kubric_train_joint.py", "joint_pose_2d3d", None),
    "train_pose3d": ("kubric_train_pose_3D.py", "gt_pose", MAKE_CKPT_POSE3D),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["demo", "train", "train_pose3d"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--iters", type=int, default=6, help="train: iterations per rank")
    ap.add_argument("--batch", type=int, default=4, help="train: objects per rank (reference: 4)")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--out", default=None, help="write a JSON summary here")
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix="forge_ref_")
    ref = stage(work)
    env = env_for(ref)
    summary = {"script": args.what, "gpus": args.gpus}
    try:
        if args.what == "demo":
            subprocess.run([sys.executable, "-c", MAKE_CKPT, "config/demo/demo.yaml"], cwd=ref, env=env, check=True)
            t0 = time.time()
            res = subprocess.run([sys.executable, "demo.py", "--cfg", "config/demo/demo.yaml"], cwd=ref, env=env,
                                 capture_output=True, text=True)
            summary.update(rc=res.returncode, seconds=round(time.time() - t0, 1),
                           gifs=sorted(os.path.basename(p) for p in glob.glob(os.path.join(ref, "output", "**", "vis_360", "*.gif"), recursive=True)))
            sys.stdout.write(res.stdout[-3000:])
            sys.stderr.write(res.stderr[-3000:])
        else:
            # the reference's yaml with a short run: resume from a synthetic epoch-1 checkpoint, stop after args.iters iterations
            script, yaml_name, ckpt_maker = TRAIN_MODES[args.what]
            src_yaml = os.path.join(ref, "config", "kubric", yaml_name + ".yaml")
            yaml_txt = open(src_yaml).read()
            import re
            yaml_txt = re.sub(r"total_iteration:\s*\d+", "total_iteration: %d" % (2 * args.iters), yaml_txt)
            yaml_txt = re.sub(r"resume:\s*\w+", "resume: True", yaml_txt)
            yaml_txt = re.sub(r"batch_size:\s*\d+", "batch_size: %d" % args.batch, yaml_txt, count=1)
            yaml_txt = re.sub(r"print_freq:\s*\d+", "print_freq: 1", yaml_txt)
            yaml_txt = re.sub(r"vis_freq:\s*\d+", "vis_freq: 1000000", yaml_txt)
            yaml_txt = re.sub(r"workers:\s*\d+", "workers: 2", yaml_txt)
            cfg = os.path.join(ref, "config", "kubric", yaml_name + "_short.yaml")
            open(cfg, "w").write(yaml_txt)
            exp = re.search(r"exp_name:\s*'?([\w-]+)'?", yaml_txt).group(1)
            out_dir = os.path.join(ref, "output", "kubric", yaml_name + "_short", exp)
            subprocess.run([sys.executable, "-c", ckpt_maker or MAKE_CKPT, cfg, out_dir], cwd=ref, env=env, check=True)
            env["FORGE_SYNTHETIC_SEQS"] = str(args.gpus * args.batch * args.iters)       # len(train_loader) = iters per rank
            launch = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                      "--master-addr", "127.0.0.1", "--master-port", "29533", "--no-python", "bash", "-c",
                      # torchrun passes --local-rank; the reference parses --local_rank
                      "exec %s %s --cfg %s --local_rank $LOCAL_RANK" % (sys.executable, script, cfg)]
            t0 = time.time()
            res = subprocess.run(launch, cwd=ref, env=env, capture_output=True, text=True)
            summary.update(rc=res.returncode, seconds=round(time.time() - t0, 1))
            lines = [l for l in (res.stdout + res.stderr).splitlines() if "Iter" in l and "Time" in l]
            summary["log_lines"] = lines[-8:]
            import re as _re
            times = [float(m.group(1)) for l in lines for m in [_re.search(r"all ([\d.]+)s", l)] if m and "rank 0" in l]
            if len(times) > 2:
                summary["s_per_step_rank0_after_warmup"] = round(sum(times[2:]) / len(times[2:]), 4)
            sys.stdout.write(res.stdout[-2500:])
            sys.stderr.write(res.stderr[-2500:])
    finally:
        if not args.keep:
            shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(summary))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(summary, fh, indent=1)
    sys.exit(0 if summary.get("rc") == 0 else 1)


if __name__ == "__main__":
    main()
