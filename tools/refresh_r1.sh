# Round-1 evidence refresh: tests, headline bench (both arms), ncu launch list, codec benches, single-stream table
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 400 gpurun_out/bench_r1.json; tail -2 gpurun_out/bench_r1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list2.log 2>&1
for w in lz4 bzip2 xz; do timeout 600 python tools/bench_codecs.py --workload $w --steps 3 --warmup 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json; done
python tools/bench_single.py 16 > gpurun_out/bench_single.json 2>&1; tail -4 gpurun_out/bench_single.json | cut -c1-200
