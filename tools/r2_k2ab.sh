cd /root/repo
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_differential.py tests/test_gpu_wrappers.py tests/test_gpu_lz4.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-legs > gpurun_out/ab_k2u.json 2> gpurun_out/ab_k2u.err || tail -3 gpurun_out/ab_k2u.err
python -c "
import json; d=json.load(open('gpurun_out/ab_k2u.json')); print('deflate', round(d['value'],1), 'GB/s', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['roofline']['kernels_ms'].items()})"
python tools/bench_codecs.py --workload lz4 --no-cpu > gpurun_out/ab_lz4u.json 2> gpurun_out/ab_lz4u.err || tail -3 gpurun_out/ab_lz4u.err
tail -1 gpurun_out/ab_lz4u.json | cut -c1-600
