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
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cat gpurun_out/${TAG}_bench_ref.json
python tools/bench_kernels.py --reps 20 --ref --bwd > gpurun_out/${TAG}_kernels.jsonl 2> gpurun_out/${TAG}_kernels.err; tail -3 gpurun_out/${TAG}_kernels.err; cat gpurun_out/${TAG}_kernels.jsonl
python tools/bench_kernels.py --reps 10 --objects 1 --img 128 --vol 32 --pts 32 --ref --only k1,k2,relayout > gpurun_out/${TAG}_kernels_cfg1.jsonl 2>/dev/null; cat gpurun_out/${TAG}_kernels_cfg1.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-clocks > gpurun_out/${TAG}_ncu_bench.log 2>&1
for K in raymarch_fwd:k1 pack_volume:relayout decoder_fwd:volrender decoder_tc:volrender rotate_fwd:k2; do
  PAT=${K%%:*}; ONLY=${K##*:}
  ncu --set full --clock-control none --import-source on -k regex:$PAT -s 2 -c 1 -f -o gpurun_out/${TAG}_$PAT \
      python tools/bench_kernels.py --reps 1 --only $ONLY > gpurun_out/${TAG}_ncu_$PAT.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
