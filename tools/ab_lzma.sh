# LZMA: parity tests and the XZ batch bench
timeout 600 python -m pytest tests/test_gpu_lzma.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/bench_codecs.py --workload xz --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('xz', d['value'], d['unit'], d['ms_per_step'])"
