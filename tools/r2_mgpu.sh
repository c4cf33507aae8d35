cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --units 131072 --distinct 2048 --steps 3 --warmup 2 --no-cpu --e2e-units 16384 > gpurun_out/r2_mgpu2.json 2> gpurun_out/r2_mgpu2.err
tail -5 gpurun_out/r2_mgpu2.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2_mgpu2.json'))
print(d['value'], d['n_gpus'], d['ms_per_step'], d['multi_gpu'], d['e2e']['value'])
PY
