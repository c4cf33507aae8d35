cd /root/repo
python -m pytest tests/test_gpu_deflate.py tests/test_gpu_differential.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu > gpurun_out/ab_k2r.json 2> gpurun_out/ab_k2r.err || tail -3 gpurun_out/ab_k2r.err
python -c "
import json; d=json.load(open('gpurun_out/ab_k2r.json')); print('K2r', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['kernels_ms'])"
SWC_DEFLATE_K2=old python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu > gpurun_out/ab_k2o.json 2> gpurun_out/ab_k2o.err
python -c "
import json; d=json.load(open('gpurun_out/ab_k2o.json')); print('K2 old', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['kernels_ms'])"
