cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/run_c5.py 2>/dev/null | tail -1 | tee gpurun_out/c5_r2_8gpu.json
C5_SPECULATE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/run_c5.py 2>/dev/null | tail -1 | tee gpurun_out/c5_r2_8gpu_spec0.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/run_c5.py 2>/dev/null | tail -1 | tee gpurun_out/c5_r2_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/run_c5.py 2>/dev/null | tail -1 | tee gpurun_out/c5_r2_2gpu.json
