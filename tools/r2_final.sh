cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_ncu_list.log 2>&1
grep -E "crc32_units|inflate_lut|lz_resolve" gpurun_out/r2_launches.csv | cut -d, -f5,15- | head -8
timeout 600 python tools/bench_gzip_multi.py 262144 > gpurun_out/r2_bench_gzip_multi.json 2>&1; tail -1 gpurun_out/r2_bench_gzip_multi.json | cut -c1-300
