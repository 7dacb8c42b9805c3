import sys, warnings, json, torch
sys.path.insert(0,'/root/repo')
from forge_b200 import synthetic as syn
from forge_b200.models.encoder import Encoder3D
from forge_b200.refine import prepare_for_pose_refinement
DEV='cuda'
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    enc=Encoder3D(syn.make_config()).to(DEV).eval()
class M: pass
m=M(); m.encoder_3d=enc; m.parameters=enc.parameters
for p in enc.parameters(): p.requires_grad_(False)
enc.channels_last_3d_(); enc.compute_dtype=torch.bfloat16
b,t,C,n=1,5,128,32
xs=[torch.randn(b,n,n,n,C,device=DEV).permute(0,4,1,2,3).requires_grad_(True) for _ in range(t)]
def run():
    for x in xs: x.grad=None
    y=enc.fuse(xs)
    d=enc.get_density3D(y); f=enc.get_render_features(y)
    (d.sum()+f.sum()).backward()
def run_fuse_only():
    for x in xs: x.grad=None
    y=enc.fuse(xs); y.sum().backward()
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    a,bb=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): g.replay()
    bb.record(); torch.cuda.synchronize()
    return a.elapsed_time(bb)/iters
out={}
for tc in (False, True):
    enc.tc_fusion=tc
    out['fuse fwd+bwd %s'%('tc' if tc else 'cudnn bf16')]=round(timed(run_fuse_only),3)
    out['fuse+heads fwd+bwd %s'%('tc' if tc else 'cudnn bf16')]=round(timed(run),3)
print(json.dumps(out))
from torch.profiler import profile, ProfilerActivity
enc.tc_fusion=True
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): run_fuse_only()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
