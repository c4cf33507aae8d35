# Late round-2 refresh (one GPU): everything the Deflate / LZ4 kernel changes touch
cd /root/repo
mkdir -p gpurun_out
VARIANTS="gpurun_ab/libswcgpu_hb1.so gpurun_ab/libswcgpu_hb16.so gpurun_ab/libswcgpu_k6.so" bash tools/r2_ab.sh
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_ncu_list.log 2>&1
timeout 600 python tools/bench_codecs.py --workload lz4 --steps 3 --warmup 2 > gpurun_out/r2_bench_lz4.json 2> gpurun_out/r2_bench_lz4.err; tail -c 200 gpurun_out/r2_bench_lz4.json; echo
timeout 600 python tools/bench_gzip_multi.py 262144 > gpurun_out/r2_bench_gzip_multi.json 2>&1; tail -1 gpurun_out/r2_bench_gzip_multi.json | cut -c1-400
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"inflate_lut|lz_resolve" -s 2 -c 2 -f -o gpurun_out/r2_prof_deflate python bench.py --steps 1 --warmup 1 --units 65536 --distinct 1024 --no-e2e --no-cpu > gpurun_out/r2_ncu_deflate.log 2>&1; tail -1 gpurun_out/r2_ncu_deflate.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"lz4_parse|lz4_exec" -s 2 -c 2 -f -o gpurun_out/r2_prof_lz4 python tools/bench_codecs.py --workload lz4 --units 65536 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_lz4.log 2>&1
