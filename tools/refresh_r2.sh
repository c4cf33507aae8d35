# Round-2 evidence refresh (one GPU): tests, headline bench (both arms), ncu launch list + full captures, codec benches,
# single-/multi-stream tables.  Everything lands in gpurun_out/; the summaries that matter are copied to profiles/ afterwards.
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_ncu_list.log 2>&1
for w in lz4 bzip2 xz; do timeout 600 python tools/bench_codecs.py --workload $w --steps 3 --warmup 2 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; tail -c 200 gpurun_out/r2_bench_$w.json; echo; done
timeout 600 python tools/bench_single.py 16 > gpurun_out/r2_bench_single.json 2>&1; tail -4 gpurun_out/r2_bench_single.json | cut -c1-220
timeout 600 python tools/bench_multi.py 512 > gpurun_out/r2_bench_multi.json 2>&1; tail -2 gpurun_out/r2_bench_multi.json | cut -c1-300
timeout 600 python tools/bench_gzip_multi.py 262144 > gpurun_out/r2_bench_gzip_multi.json 2>&1; tail -1 gpurun_out/r2_bench_gzip_multi.json | cut -c1-400
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"inflate_lut|lz_resolve" -s 2 -c 2 -f -o gpurun_out/r2_prof_deflate python bench.py --steps 1 --warmup 1 --units 65536 --distinct 1024 --no-e2e --no-cpu > gpurun_out/r2_ncu_deflate.log 2>&1; tail -1 gpurun_out/r2_ncu_deflate.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"lz4_parse|lz4_exec" -s 2 -c 2 -f -o gpurun_out/r2_prof_lz4 python tools/bench_codecs.py --workload lz4 --units 65536 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_lz4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"bzip2_kernel" -s 1 -c 1 -f -o gpurun_out/r2_prof_bzip2 python tools/bench_codecs.py --workload bzip2 --units 1024 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_bzip2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"lzma_kernel" -s 1 -c 1 -f -o gpurun_out/r2_prof_lzma python tools/bench_codecs.py --workload xz --units 592 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_lzma.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
