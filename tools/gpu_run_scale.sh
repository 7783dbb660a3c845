cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
tail -c 300 gpurun_out/bench_r2_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_${N}gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms'], d['gpu_launches'], d['best_node'])
print(json.dumps(d.get('per_rank')))
PY
