cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r4c_bench8.json 2> gpurun_out/r4c_bench8.err; echo bench8 rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r4c_bench4.json 2> gpurun_out/r4c_bench4.err; echo bench4 rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r4c_bench2.json 2> gpurun_out/r4c_bench2.err; echo bench2 rc=$?
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
