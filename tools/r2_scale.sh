cd /root/repo
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_differential.py -m gpu -x -q 2>&1 | tail -2
for v in p31 p3 p1; do
for u in 151552 262144; do
  SWCGPU_SO=/root/repo/gpurun_ab/libswcgpu_$v.so python bench.py --units $u --distinct 1024 --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs > gpurun_out/sc_$u.json 2> gpurun_out/sc_$u.err || tail -3 gpurun_out/sc_$u.err
  python -c "
import json; d=json.load(open('gpurun_out/sc_$u.json')); k=d['roofline']['kernels_ms']; print('$v', $u, {a: round(v,2) for a,v in k.items()})"
done; done
