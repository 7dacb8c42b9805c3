#!/bin/bash
# One GPU visit: parity tests, bench line, per-kernel timings, ncu launch list + full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [pytest-target]
TAG=${1:-r01}
TARGET=${2:-tests}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest $TARGET -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
python tools/bench_kernels.py --reps 20 --ref > gpurun_out/${TAG}_kernels.jsonl 2> gpurun_out/${TAG}_kernels.err; tail -3 gpurun_out/${TAG}_kernels.err; cat gpurun_out/${TAG}_kernels.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raymarch_fwd -s 2 -c 1 -f -o gpurun_out/${TAG}_k1 \
    python tools/bench_kernels.py --reps 1 --only k1 > gpurun_out/${TAG}_ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rotate_fwd -s 2 -c 1 -f -o gpurun_out/${TAG}_k2 \
    python tools/bench_kernels.py --reps 1 --only k2 > gpurun_out/${TAG}_ncu_k2.log 2>&1
ls -la gpurun_out | tail -20
