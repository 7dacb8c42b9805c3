set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FORGE_K1T_RING=11 timeout 300 python -m pytest tests/test_gpu_k1_tma.py tests/test_gpu_forward.py -x -q -m gpu > gpurun_out/r02k_pytest_ring11.log 2>&1; echo "pytest ring11 rc=$?"
tail -n 5 gpurun_out/r02k_pytest_ring11.log
(
for ring in 10 11; do FORGE_K1T_RING=$ring python tools/bench_kernels.py --only k1 --reps 40 | sed "s/^/ring=$ring /"; done
for ring in 10 11; do FORGE_K1T_RING=$ring python tools/bench_kernels.py --only k1 --reps 10 --objects 8 --img 512 --vol 128 --pts 128 | sed "s/^/cfg4 ring=$ring /"; done
for shape in 6 0; do for stream in 0 1; do FORGE_K2_SHAPE=$shape FORGE_K2_STREAM=$stream python tools/bench_kernels.py --only k2 --reps 40 | grep "transform jobs"; done; done
FORGE_K2_LEAN=0 python tools/bench_kernels.py --only k2 --reps 40 | grep "transform jobs"
) > gpurun_out/r02k_ab.jsonl 2>&1
cat gpurun_out/r02k_ab.jsonl
