set -x
cd /root/repo
K=${KERNEL:-inflate_lut_kernel}; O=${OUT:-r2_prof10}
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/$O -f python bench.py --units 65536 --distinct 1024 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs > gpurun_out/$O.log 2>&1
tail -3 gpurun_out/$O.log | cut -c1-300
