cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log
tail -6 gpurun_out/r2k_tests.log
python profiles/dbg_circ.py 2>&1 | tail -6
python profiles/r2_configs.py pct 2>&1 | tail -2
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2k_bench.err
