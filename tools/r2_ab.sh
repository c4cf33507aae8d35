# A/B of library variants under gpurun_ab/ (and the baseline copy at the repo root) on the headline workload
cd /root/repo
for so in ${VARIANTS:-libswcgpu_base.so gpurun_ab/*.so}; do
  name=$(basename $so .so)
  SWCGPU_SO=/root/repo/$so timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-legs > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/ab_$name.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/ab_$name.json')); print('$name', round(d['value'],1), 'GB/s', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['roofline']['kernels_ms'].items()})"
done
