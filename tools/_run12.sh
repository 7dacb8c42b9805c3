cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for bm in 0 1; do
python - <<PY 2>/dev/null
import json, torch, argparse
torch.backends.cudnn.benchmark = bool($bm)
import bench
dev = torch.device('cuda:0')
r = bench.e2e_model_leg(dev, argparse.Namespace())
print('cudnn.benchmark=$bm', {k: round(v['ms_per_step'],3) for k,v in r.items() if isinstance(v, dict)})
PY
done
python tools/model_bench.py --no-ref 2>/dev/null | tail -n 3
