set -x
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:lz_resolve_kernel -s 1 -c 1 -o gpurun_out/r2_prof11 -f python bench.py --units 65536 --distinct 1024 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs > gpurun_out/r2_prof11.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_exec_kernel -s 1 -c 1 -o gpurun_out/r2_prof12 -f python tools/bench_codecs.py --workload lz4 --units 65536 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_prof12.log 2>&1
tail -2 gpurun_out/r2_prof12.log | cut -c1-200
