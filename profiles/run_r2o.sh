cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2o_bench_n2.err
head -c 300 gpurun_out/r2o_bench_n2.json
