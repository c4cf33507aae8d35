# A/B: K2 overlapped with K1 (default) vs serial; then parity tests in the default mode
mkdir -p gpurun_out
for mode in serial overlap; do
  echo "SWC_DEFLATE_K2=$mode"
  SWC_DEFLATE_K2=$mode timeout 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$mode.log 2>&1
  echo "rc=$?"; tail -c 1500 gpurun_out/ab_$mode.log
done
timeout 300 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -2
