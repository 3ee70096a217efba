cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log
tail -25 gpurun_out/r2g_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2g_bench.err
head -c 6000 gpurun_out/r2g_bench.json
