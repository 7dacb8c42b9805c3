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


def main():
    from oracle import enable_p3d_standin
    enable_p3d_standin()
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    golden_volrender("volrender_small", n_obj=1, n_views=3, img_size=32, vol=12, n_pts=24, dense=False, seed=11)
    golden_volrender("volrender_dense", n_obj=2, n_views=2, img_size=32, vol=10, n_pts=16, dense=True, seed=12)
    golden_rotate("rotate_g16", n_obj=1, n_views=3, channels=4, grid=16, seed=21)
    golden_rotate("rotate_g32", n_obj=1, n_views=2, channels=2, grid=32, seed=22)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
