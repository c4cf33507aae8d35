for cfg in "" "-DSWC_ACC32" "-DSWC_REFILL_PRED" ; do
  (cd swcompression_b200/csrc && touch inflate.cu && make -j8 EXTRA="$cfg" > /dev/null 2>&1)
  echo "cfg=[$cfg]"; python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernels_ms'])"
done
