cd /root/repo
python -m pytest tests/test_gpu_lz4.py tests/test_gpu_deflate.py -m gpu -x -q 2>&1 | tail -2
for c in 2 4 6; do
  SWC_K2_CTAS=$c python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs > gpurun_out/ab_res$c.json 2> gpurun_out/ab_res$c.err || tail -3 gpurun_out/ab_res$c.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_res$c.json')); print('K2 ctas $c', {k: round(v,2) for k,v in d['roofline']['kernels_ms'].items()})"
done
for v in l6 l8; do
  SWCGPU_SO=/root/repo/gpurun_ab/libswcgpu_$v.so python tools/bench_codecs.py --workload lz4 --no-cpu --steps 2 --warmup 1 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('lz4 $v', d['roofline']['intervals_between_timing_marks_ms'][:3])"
done
