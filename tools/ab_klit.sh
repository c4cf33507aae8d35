# A/B of K1's literal steps per round (after moving the length-symbol tail into the match phase)
for cfg in "-DSWC_KLIT=4" "-DSWC_KLIT=3" "-DSWC_KLIT=6"; do
  (cd swcompression_b200/csrc && touch inflate.cu && make -j8 EXTRA="$cfg" > /dev/null 2>&1)
  echo "cfg=[$cfg]"
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms'])"
done
(cd swcompression_b200/csrc && touch inflate.cu && make -j8 > /dev/null 2>&1)
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -2
