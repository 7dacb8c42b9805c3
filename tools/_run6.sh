cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_forward.py -x -q -m gpu > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02o_pytest.log
(
for occ in 4 5 6; do FORGE_K2_OCC=$occ python tools/bench_kernels.py --only k2 --reps 40 | grep "transform jobs" | sed "s/^/occ=$occ /"; done
for occ in 4 6; do FORGE_K2_OCC=$occ python tools/bench_kernels.py --only k2 --reps 10 --objects 8 --img 512 --vol 128 --pts 128 | grep "transform jobs" | sed "s/^/cfg4 occ=$occ /"; done
python tools/bench_kernels.py --bwd --only none --reps 20 | grep rotate
) > gpurun_out/r02o_ab.jsonl 2>&1
cat gpurun_out/r02o_ab.jsonl
