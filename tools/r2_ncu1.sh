set -x
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:'lz_replay_kernel' -s 1 -c 1 -o gpurun_out/r2_prof9 -f python bench.py --units 65536 --distinct 1024 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2_ncu9.log 2>&1
tail -3 gpurun_out/r2_ncu9.log | cut -c1-300
