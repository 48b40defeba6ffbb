# 8-GPU runs kept under profiles/: the default scaling line (G1, 256 per GPU, + configs[3] 512 per GPU) and configs[4]
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 "$@"; }
timeout 600 env tr_dummy=1 bash -c 'true'
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_8gpu_g1.json 2> gpurun_out/bench_8gpu_g1.err ); echo "g1 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_8gpu_g1.json'))
    print('G1 x8:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['runs_ms_per_step'], 'cfg3', d.get('configs3_g1_batch4096_8gpu'))
except Exception as e:
    print('g1 parse failed', e)
PY
tail -3 gpurun_out/bench_8gpu_g1.err | cut -c1-300
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --workload stress --global-batch 8192 --steps 3 --warmup 3 --state-warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_stress.json 2> gpurun_out/bench_8gpu_stress.err ); echo "stress rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_8gpu_stress.json'))
    print('stress x8:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['runs_ms_per_step'], d['config'])
except Exception as e:
    print('stress parse failed', e)
PY
tail -4 gpurun_out/bench_8gpu_stress.err | cut -c1-300
