mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool: K1-T (TMA bricks) small cases + conv3d_tc small cases" >> gpurun_out/r02_sanitizer.txt
  timeout 500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_k1_tma.py -q -x -k "64-16-24 or 32-10-16 or missing or without_depth" 2>&1 | tail -6 >> gpurun_out/r02_sanitizer.txt
  timeout 500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_conv3d_tc.py -q -x -k "1-4-4-8-64 or 2-8-8-8 or gru_gate or strided" 2>&1 | tail -6 >> gpurun_out/r02_sanitizer.txt
done
cat gpurun_out/r02_sanitizer.txt
