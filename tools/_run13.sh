cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_model.py tests/test_gpu_host_rows.py -x -q -m gpu > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/r02u_pytest.log
(
for g in 0 1; do FORGE_K2B_GATHER=$g python tools/bench_kernels.py --bwd --only none --reps 20 | grep rotate | sed "s/^/gather=$g /"; done
for g in 0 1; do FORGE_K2B_GATHER=$g python tools/bench_kernels.py --bwd --only none --reps 5 --objects 8 --img 512 --vol 128 --pts 128 | grep rotate | sed "s/^/cfg4 gather=$g /"; done
) > gpurun_out/r02u_ab.jsonl 2>&1
cat gpurun_out/r02u_ab.jsonl
