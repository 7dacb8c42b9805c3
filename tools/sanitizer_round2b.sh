# compute-sanitizer over the kernels added / changed in the second half of round 2 (small cases): K1 backward with the same-base
# merge (shared-memory exchange ordered by __syncwarp), quad density gradient + fold kernel, lean K2 forward, shared-memory-free pack,
# tuned K1-T loop.  Usage (under gpurun): bash tools/sanitizer_round2b.sh
mkdir -p gpurun_out
OUT=gpurun_out/r02b_sanitizer.txt
: > $OUT
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool: K1 backward (merge + quads), K2 lean, pack, K1-T tuned" >> $OUT
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_backward.py -q -x -k "vs_oracle and (False or True) and not merge_patterns and not rotate or non_cubic" 2>&1 | tail -6 >> $OUT
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_forward.py -q -x -k "rotate_vs_oracle and (1-5-128-16 or 1-2-6-20) or relayout" 2>&1 | tail -6 >> $OUT
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_k1_tma.py -q -x -k "64-16-24 or 32-10-16 or missing" 2>&1 | tail -6 >> $OUT
done
cat $OUT
