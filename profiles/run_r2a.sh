set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 600 python profiles/r2_configs.py circ square > gpurun_out/r2a_configs.jsonl 2> gpurun_out/r2a_configs.err
cat gpurun_out/r2a_configs.jsonl; tail -5 gpurun_out/r2a_configs.err
