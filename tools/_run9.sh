cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r02r_pytest.log
python - > gpurun_out/r02r_e2e_model.json 2> gpurun_out/r02r_e2e_model.err <<'PY'
import json, torch, argparse
import bench
dev = torch.device('cuda:0')
print(json.dumps(bench.e2e_model_leg(dev, argparse.Namespace())))
PY
tail -n 5 gpurun_out/r02r_e2e_model.err; cat gpurun_out/r02r_e2e_model.json
