#!/bin/bash
# ncu capture of the K1/K2 backward kernels. Usage (under gpurun): bash tools/ncu_bwd.sh <tag>
TAG=${1:-bwd}
mkdir -p gpurun_out
python tools/bench_kernels.py --reps 10 --bwd --only none
ncu --set full --clock-control none --import-source on -k regex:raymarch_bwd -s 2 -c 1 -f -o gpurun_out/${TAG}_k1bwd \
    python tools/bench_kernels.py --reps 1 --bwd --only none > gpurun_out/${TAG}_ncu_k1bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rotate_bwd -s 2 -c 1 -f -o gpurun_out/${TAG}_k2bwd \
    python tools/bench_kernels.py --reps 1 --bwd --only none > gpurun_out/${TAG}_ncu_k2bwd.log 2>&1
ls -la gpurun_out | grep ${TAG}
