cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=r02n
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -n 2 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -n 3 gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; cat gpurun_out/${T}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clocks --quick > gpurun_out/${T}_ncu_bench.log 2>&1
python tools/bench_kernels.py --reps 20 --bwd --only none > gpurun_out/${T}_kernels_bwd.jsonl 2>&1; cat gpurun_out/${T}_kernels_bwd.jsonl
python tools/refine_bench.py --no-ref > gpurun_out/${T}_refine.json 2>gpurun_out/${T}_refine.err; tail -n 12 gpurun_out/${T}_refine.json
