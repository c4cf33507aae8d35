# A/B of the warp-per-unit decoder's window size (single-stream latency + small-batch throughput)
for cfg in "-DSWC_WIN_WORDS=9" "-DSWC_WIN_WORDS=19" ; do
  (cd swcompression_b200/csrc && touch inflate_warp.cu && make -j8 EXTRA="$cfg" > /dev/null 2>&1)
  echo "cfg=[$cfg]"
  SWC_DEFLATE_K1=warp python bench.py --steps 3 --warmup 2 --units 16384 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('16384 units', d['value'], d['roofline']['kernels_ms'])"
  python - <<'PY'
import sys, time, zlib
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import helpers as H
from swcompression_b200 import Deflate, GzipArchive
raw = b"".join(H.textlike(65536, 9000+i) for i in range(64))   # 4 MiB, one multi-block stream
comp = H.raw_deflate(raw)
Deflate.decompress(comp)
t=time.perf_counter(); out = Deflate.decompress(comp); dt=time.perf_counter()-t
assert out == raw
print('single 4 MiB stream: %.1f ms -> %.1f MB/s decompressed' % (dt*1e3, len(raw)/dt/1e6))
PY
done
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_wrappers.py -x -q 2>&1 | tail -2
