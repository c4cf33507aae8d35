# Round-1 evidence refresh: tests, headline bench (both arms), ncu launch list + one full capture, codec benches, single-/multi-stream tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 400 gpurun_out/bench_r1.json; tail -2 gpurun_out/bench_r1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list2.log 2>&1
for w in lz4 bzip2 xz; do timeout 600 python tools/bench_codecs.py --workload $w --steps 3 --warmup 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 300 gpurun_out/bench_$w.json; done
timeout 600 python tools/bench_single.py 16 > gpurun_out/bench_single.json 2>&1; tail -4 gpurun_out/bench_single.json | cut -c1-200
timeout 600 python tools/bench_multi.py 512 > gpurun_out/bench_multi.json 2>&1; tail -2 gpurun_out/bench_multi.json | cut -c1-300
timeout 600 python tools/bench_gzip_multi.py 262144 > gpurun_out/bench_gzip_multi.json 2>&1; tail -1 gpurun_out/bench_gzip_multi.json | cut -c1-400
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"inflate_huffman|lz_resolve" -c 2 -o gpurun_out/prof_r1_final2 python bench.py --steps 1 --warmup 0 --units 65536 --no-e2e --no-cpu > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
