cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3g_bench8.json 2> gpurun_out/r3g_bench8.err; echo bench8 rc=$?
grep "^\[bench" gpurun_out/r3g_bench8.err | tail -8
