# bzip2: per-phase cycle counts of one 900 KB block (SWC_BZ_PROFILE build), then parity tests and the batch bench (default build)
bash tools/prof_bz2.sh 2>&1 | tail -4
(cd swcompression_b200/csrc && touch bzip2.cu && make -j8 > /dev/null 2>&1)
timeout 600 python -m pytest tests/test_gpu_bzip2.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/bench_codecs.py --workload bzip2 --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bzip2', d['value'], d['unit'], d['ms_per_step'])"
