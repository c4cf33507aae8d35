# A/B of the warp-per-unit decoder's literal steps per round; device-resident 16384-unit batch + e2e (host buffers, 65536 units)
mkdir -p gpurun_out
for cfg in "-DSWC_KW_LIT=4" "-DSWC_KW_LIT=2" "-DSWC_KW_LIT=6"; do
  (cd swcompression_b200/csrc && touch inflate_warp.cu && make -j8 EXTRA="$cfg" > /dev/null 2>&1)
  echo "cfg=[$cfg]"
  timeout 200 python bench.py --steps 5 --warmup 3 --units 16384 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('16384 units', d['value'], d['roofline']['kernels_ms'])"
  timeout 200 python bench.py --steps 1 --warmup 3 --units 65536 --no-cpu --e2e-steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e'])"
done
(cd swcompression_b200/csrc && touch inflate_warp.cu && make -j8 > /dev/null 2>&1)
timeout 500 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -2
