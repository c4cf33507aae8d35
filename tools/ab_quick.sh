# quick check: config-2 bench (device-resident) + Deflate/wrapper parity tests
mkdir -p gpurun_out
timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/ab.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms'])"
timeout 400 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_wrappers.py -x -q -m gpu 2>&1 | tail -2
