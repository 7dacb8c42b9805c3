"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden

It puts the PyTorch3D stand-in (oracle/p3d_standin) and /root/reference on sys.path, imports
``models.volume_render.VolRender`` and ``models.rotate.Rotate_world`` exactly as the reference
scripts do, feeds them seeded inputs, and stores inputs + outputs.  It also asserts that
``oracle/reference_path.py`` (the restated call sequence that travels to the GPU box)
reproduces the reference modules bit-for-bit on this machine.

What the fixtures pin: the reference's glue code (K halving, tuple ordering, permutes, decoder,
upsampling, T = pose_0 pose_1^-1, align_corners=False resample, view-0 passthrough).  What they
cannot pin: PyTorch3D-internal arithmetic (stand-in) -> "parity unpinned" for that part.
"""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FORGE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def _cfg(img_size, n_pts, min_depth=0.5, max_depth=2.0, volume_size=1.0):
    return SimpleNamespace(dataset=SimpleNamespace(img_size=img_size, name='kubric'),
                           render=SimpleNamespace(n_pts_per_ray=n_pts, min_depth=min_depth, max_depth=max_depth,
                                                  volume_size=volume_size, k_size=5, camera_z=1.5),
                           network=SimpleNamespace(padding_mode='zeros'))


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def golden_volrender(name, n_obj, n_views, img_size, vol, n_pts, dense, seed):
    from models.volume_render import VolRender           # the reference module, unmodified
    from oracle import reference_path as rp
    from forge_b200 import synthetic as syn
    cfg = _cfg(img_size, n_pts)
    torch.manual_seed(seed)
    ref = VolRender(cfg)
    # non-trivial BN statistics so eval-mode BN is exercised
    for m in ref.conv_rgb:
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    ref.eval()
    inp = syn.render_inputs(n_obj, n_views, img_size, vol, seed=seed, dense=dense)
    feat_all = inp['feat'][inp['view2vol'].long()]       # as-called: one volume per view (models/model.py:138-139)
    dens_all = inp['dens'][inp['view2vol'].long()]
    cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb, sil, depth, oproj = ref(cam, feat_all, dens_all, render_depth=True, return_origin_proj=True)
        # the 18-channel image straight out of the renderer (before the decoder)
        cam2 = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
        core, _ = rp.raymarch_core(ref.renderer, cam2, feat_all, dens_all, img_size, cfg.render.volume_size, True)
        # restated call sequence must equal the reference module
        cam3 = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
        o_rgb, o_sil, o_depth, o_oproj = rp.volrender_forward(ref.renderer, ref.conv_rgb, cam3, feat_all, dens_all,
                                                              img_size, cfg.render.volume_size, True, True)
    # the reference mutates the caller's K in place (models/volume_render.py:50-51)
    assert torch.allclose(cam['K'][:, 0, 0], inp['K'][:, 0, 0] / 2) and (cam['K'][:, 2, 2] == 1).all()
    for a, b in ((rgb, o_rgb), (sil, o_sil), (depth, o_depth), (oproj, o_oproj)):
        assert torch.equal(a, b), "oracle/reference_path.py diverged from the reference module"
    sd = {"sd." + k: v for k, v in ref.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        feat=inp['feat'], dens=inp['dens'], view2vol=inp['view2vol'], R=inp['R'], T=inp['T'], K=inp['K'],
        img_size=img_size, n_pts=n_pts, min_depth=cfg.render.min_depth, max_depth=cfg.render.max_depth,
        volume_size=cfg.render.volume_size, core=core, rgb=rgb, sil=sil, depth=depth, origin_proj=oproj, **sd)))
    print(name, "core", tuple(core.shape), "rgb", tuple(rgb.shape), "opacity max", core[..., 16].max().item())


def golden_rotate(name, n_obj, n_views, channels, grid, seed):
    from models.rotate import Rotate_world               # the reference module, unmodified
    from oracle import reference_path as rp
    from forge_b200 import synthetic as syn
    torch.manual_seed(seed)
    ref = Rotate_world(_cfg(256, 64))
    vox, poses = syn.rotate_inputs(n_obj, n_views, channels, grid, seed=seed)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = ref(vox, poses, grid_size=grid)
        o_out = rp.rotate_world_forward(vox, poses, grid, 1.0)
    assert torch.equal(out, o_out), "oracle/reference_path.py diverged from the reference module"
    gmax = {32: ref.grid_coord_max, 16: ref.grid_coord_max_16, 64: ref.grid_coord_max_64,
            128: ref.grid_coord_max_128, 48: ref.grid_coord_max_48}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        voxels=vox, poses=poses, grid=grid, out=out, T=ref.get_transformation(poses),
        grid_coord_max=np.array([gmax[k] for k in (16, 32, 48, 64, 128)], dtype=np.float64))))
    print(name, tuple(out.shape), "grid max", gmax[grid])


def _offline_torchvision():
    """The reference constructs torchvision ResNet-50 / VGG-16 with ``pretrained=True`` (models/encoder.py:72): there is
    no network here, and the values are overwritten by the seeded state_dict anyway -> build them with random init."""
    import torchvision
    for name in ('resnet50', 'vgg16'):
        orig = getattr(torchvision.models, name)
        if getattr(orig, '_forge_offline', False):
            continue

        def offline(*a, _orig=orig, **kw):
            kw.pop('pretrained', None)
            kw['weights'] = None
            return _orig(*a, **kw)
        offline._forge_offline = True
        setattr(torchvision.models, name, offline)


def golden_encoder(name, seed):
    """Reference Encoder3D (models/encoder.py:46-63) and ConvGRU_3D (models/fusion.py:71-95), eval mode, seeded weights."""
    from models.encoder import Encoder3D                  # the reference module, unmodified
    from oracle import seeded
    torch.manual_seed(seed)
    ref = seeded.load_seeded(Encoder3D(_cfg(256, 64)), seed).eval()
    img = seeded.seeded_tensor(seed, 'img', (1, 3, 32, 32), kind='rand')
    views = seeded.seeded_tensor(seed, 'views', (1, 3, 128, 8, 8, 8))
    with torch.no_grad():
        feat3d = ref.get_feat3D(img)                      # [1,128,32,4,4]
        fused = ref.fuse(views)                           # [1,128,8,8,8]
        dens = ref.get_density3D(fused)                   # [1,1,16,16,16]
        rfeat = ref.get_render_features(fused)            # [1,16,16,16,16]
        # the cell alone, one step from a given state (models/fusion.py:21-35)
        h0 = seeded.seeded_tensor(seed, 'h0', (1, 128, 8, 8, 8))
        cell = ref.fusion_feature.cells[0](views[:, 0], h0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        seed=seed, feat3d=feat3d, fused=fused, dens=dens, rfeat=rfeat, cell=cell)))
    print(name, "feat3d", tuple(feat3d.shape), float(feat3d.abs().max()), "fused", float(fused.abs().max()),
          "dens>0", float((dens > 0).float().mean()))


def golden_convgru_train(name, seed):
    """ConvGRU_3D in TRAIN mode (batch-statistics BatchNorm3d in fusion_conv / fusion_norm, models/fusion.py:60-68):
    output and the updated running statistics."""
    from models.fusion import ConvGRU_3D                  # the reference module, unmodified
    from oracle import seeded
    ref = seeded.load_seeded(ConvGRU_3D(_cfg(256, 64), n_layers=1, input_size=16, hidden_size=16), seed).train()
    x = seeded.seeded_tensor(seed, 'x', (2, 3, 16, 6, 6, 6))
    out = ref(x, [ref.fusion_conv(x.mean(dim=1))])
    out.square().sum().backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        seed=seed, out=out, norm_mean=ref.fusion_norm.running_mean, norm_var=ref.fusion_norm.running_var,
        conv_mean=ref.fusion_conv[1].running_mean, g_gate=ref.cells[0].conv_gate.weight.grad,
        g_fconv=ref.fusion_conv[0].weight.grad)))
    print(name, tuple(out.shape), float(out.detach().abs().max()))


def golden_pose3d_model(name, seed, train_heads):
    """Reference FORGE_poseEstimator3D.forward (models/model_single_pose_estimator.py:26-138) with ground-truth poses:
    the view -> volume order (2-view volume x3, 3-view volume x2, all-view volume x5) and the head batching."""
    from models.model_single_pose_estimator import FORGE_poseEstimator3D      # the reference module, unmodified
    from oracle import seeded
    from forge_b200 import synthetic as syn
    cfg = _cfg(256, 32)
    cfg.train = SimpleNamespace(use_gt_pose=True, canonicalize=True, parameter='all')
    cfg.network.rot_representation = 'quat'
    torch.manual_seed(seed)
    ref = seeded.load_seeded(FORGE_poseEstimator3D(cfg), seed).eval()
    if train_heads:           # the sub-modules kubric_train_pose_3D.py:38-43 puts in train mode, plus the heads
        ref.encoder_3d.fusion_feature.train()
        ref.encoder_3d.density_head.train()
        ref.encoder_3d.features_head.train()
        ref.render.train()
    sample = syn.kubric_batch(2 if train_heads else 1, n_views_all=5, img_size=256, seed=seed)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb, mask = ref(sample, None, 'cpu')
    extra = {}
    if train_heads:
        extra = dict(dens_bn_mean=ref.encoder_3d.density_head[1].running_mean,
                     feat_bn_var=ref.encoder_3d.features_head[4].running_var,
                     rgb_bn_mean=ref.render.conv_rgb[1].running_mean)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        seed=seed, rgb_sub=seeded.subsample(rgb), mask_sub=seeded.subsample(mask), rgb_mean=rgb.mean(dim=(1, 2, 3)),
        mask_mean=mask.mean(dim=(1, 2, 3)), **extra)))
    print(name, tuple(rgb.shape), "rgb max", float(rgb.max()), "mask max", float(mask.max()))


def golden_joint_model(name, seed):
    """Reference FORGE.forward (models/model.py:42-148) with PREDICTED poses: both pose networks + pose head -> canonical pose
    algebra -> rotate -> distance-sorted fuse -> heads -> 5 input + 2 novel views."""
    from models.model import FORGE                        # the reference module, unmodified
    from oracle import seeded
    from forge_b200 import synthetic as syn
    cfg = _cfg(256, 32)
    cfg.train = SimpleNamespace(use_gt_pose=False, canonicalize=True, parameter='joint')
    cfg.network.rot_representation = 'quat'
    torch.manual_seed(seed)
    ref = seeded.load_seeded(FORGE(cfg), seed).eval()
    ref.encoder_3d.density_head[6].bias.data.fill_(0.1)
    sample = syn.kubric_batch(1, n_views_all=7, img_size=256, seed=seed)

    class _DS:      # the two dataset helpers forward() calls (reference dataset/kubric.py:448-452)
        ext = torch.eye(4)
        ext[2, 3] = cfg.render.camera_z

        def get_canonical_extrinsics_cv2(self, device='cpu'):
            return self.ext.to(device)

        def get_canonical_pose_cv2(self, device='cpu'):
            return torch.inverse(self.ext).to(device)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb, mask, oproj, poses = ref(sample, _DS(), 'cpu')
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(
        seed=seed, rgb_sub=seeded.subsample(rgb), mask_sub=seeded.subsample(mask), rgb_mean=rgb.mean(dim=(1, 2, 3)),
        mask_mean=mask.mean(dim=(1, 2, 3)), origin_proj=oproj, pose_pred=poses['pred'], pose_gt=poses['gt'], pose_conf=poses['conf'])))
    print(name, tuple(rgb.shape), "mask max", float(mask.max()), "pred", poses['pred'][0].tolist())


def golden_mat2quat(name, seed):
    """utils/geo_utils.mat2quat (:140-207) including rotations by exactly / nearly 180 degrees."""
    from utils import geo_utils                           # the reference module, unmodified
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(64, 4, generator=g)
    axes = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 1, 1], [0.3, -0.5, 0.8], [-1, 0.2, 0]])
    axes = axes / axes.norm(dim=1, keepdim=True)
    half_turn = torch.cat([torch.zeros(7, 1), axes], dim=1)
    near = torch.cat([torch.full((7, 1), 1e-4), axes], dim=1)
    q = torch.cat([q, half_turn, near])
    q = q / q.norm(dim=1, keepdim=True)
    T = torch.eye(4).repeat(q.shape[0], 1, 1)
    T[:, :3, :3] = geo_utils.quat2mat_transform(q)
    T[:, :3, 3] = torch.randn(q.shape[0], 3, generator=g)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(dict(T=T, quat7=geo_utils.mat2quat(T))))
    print(name, tuple(T.shape))


def main(which=None):
    from oracle import enable_p3d_standin
    enable_p3d_standin()
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    _offline_torchvision()
    # the reference's pose_estimator_2d.py:96 fetches ResNet weights through model_zoo: reuse compat's offline fallback
    import importlib.util
    spec = importlib.util.spec_from_file_location("forge_compat_site", os.path.join(ROOT, "compat", "sitecustomize.py"))
    spec.loader.exec_module(importlib.util.module_from_spec(spec))
    if not hasattr(np, 'float'):          # the reference's models/model_utils.py:45 still uses the removed alias
        np.float = float
    jobs = dict(
        volrender_small=lambda: golden_volrender("volrender_small", n_obj=1, n_views=3, img_size=32, vol=12, n_pts=24,
                                                 dense=False, seed=11),
        volrender_dense=lambda: golden_volrender("volrender_dense", n_obj=2, n_views=2, img_size=32, vol=10, n_pts=16,
                                                 dense=True, seed=12),
        rotate_g16=lambda: golden_rotate("rotate_g16", n_obj=1, n_views=3, channels=4, grid=16, seed=21),
        rotate_g32=lambda: golden_rotate("rotate_g32", n_obj=1, n_views=2, channels=2, grid=32, seed=22),
        encoder_small=lambda: golden_encoder("encoder_small", seed=31),
        convgru_train=lambda: golden_convgru_train("convgru_train", seed=32),
        pose3d_model_eval=lambda: golden_pose3d_model("pose3d_model_eval", seed=33, train_heads=False),
        pose3d_model_train=lambda: golden_pose3d_model("pose3d_model_train", seed=34, train_heads=True),
        mat2quat=lambda: golden_mat2quat("mat2quat", seed=35),
        joint_model_eval=lambda: golden_joint_model("joint_model_eval", seed=36),
    )
    for k in (which or jobs):
        jobs[k]()


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main(sys.argv[1:] or None)
