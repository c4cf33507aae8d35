cd /root/repo
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | tail -3
tail -2 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms'], 'path_frac', d['roofline']['path_frac'])
print('e2e', d['e2e']); print('cpu', d['cpu_baseline']); print(d['clocks'], d['gpu_launches'])
PY
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json ) 2>&1 | tail -3
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ref.json')); print('ref', d['value'], d['cpu_baseline'])"
