set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py -x -q > gpurun_out/r2b_box.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_box.log
tail -30 gpurun_out/r2b_box.log
timeout 300 python profiles/r2_configs.py square > gpurun_out/r2b_configs.jsonl 2> gpurun_out/r2b_configs.err
cat gpurun_out/r2b_configs.jsonl; tail -5 gpurun_out/r2b_configs.err
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_box.py > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
