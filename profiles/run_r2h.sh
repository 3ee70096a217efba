cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2h_bench.err
head -c 3000 gpurun_out/r2h_bench.json
timeout 300 python -m pytest tests/test_gpu_plugins.py -q -x 2>&1 | tail -5
