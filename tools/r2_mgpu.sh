cd /root/repo
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_mgpu2.json 2> gpurun_out/r2_mgpu2.err ) 2>&1 | tail -3
wc -l gpurun_out/r2_mgpu2.json; tail -3 gpurun_out/r2_mgpu2.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_mgpu2.json'))
print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('units_per_step'))
print(json.dumps(d['multi_gpu'])[:900])
PY
