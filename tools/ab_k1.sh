# A/B builds of K1 on the GPU box (nvcc is in the image): prints decompressed GB/s and per-kernel ms for each variant
for cfg in "-DSWC_K1_UNIFIED=2" "-DSWC_K1_UNIFIED=4" "-DSWC_K1_UNIFIED=8" ; do
  (cd swcompression_b200/csrc && touch inflate.cu && make -j8 EXTRA="$cfg" > /dev/null 2>&1)
  echo "cfg=[$cfg]"; python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernels_ms'])"
done
python -m pytest tests/test_gpu_deflate.py -x -q 2>&1 | tail -2
