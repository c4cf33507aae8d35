set -x
cd /root/repo
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_differential.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; tail -3 gpurun_out/r2_b1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b1.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms'], d['roofline']['path_frac'])"
