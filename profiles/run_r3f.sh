cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3f_bench8.json 2> gpurun_out/r3f_bench8.err; echo bench8 rc=$?
tail -c 1500 gpurun_out/r3f_bench8.err | grep bench
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 --no-configs > gpurun_out/r3f_bench4.json 2> gpurun_out/r3f_bench4.err; echo bench4 rc=$?
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
