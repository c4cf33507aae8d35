cd /root/repo
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_differential.py -m gpu -x -q 2>&1 | tail -3
for v in default G H I J; do
  so=/root/repo/gpurun_ab/lib_$v.so; [ $v = default ] && so=/root/repo/swcompression_b200/libswcgpu.so
  SWCGPU_SO=$so python bench.py --units 131072 --distinct 2048 --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', round(d['value'],1), d['roofline']['kernels_ms'])"
done
