set -x
cd $GRAFT_REPO_ROOT
python profiles/sweep_box.py "NBH_BOX_WAIT_NS=0" "NBH_BOX_WAIT_NS=1000" "NBH_BOX_WAIT_NS=100" "NBH_BOX_WAIT_NS=0 NBH_BOX_PER_SM=1" "NBH_BOX_WAIT_NS=0 NBH_BOX_STAGES=2" "NBH_SQ_BOX=0" > gpurun_out/r2d_sweep.jsonl 2>&1
cat gpurun_out/r2d_sweep.jsonl
timeout 300 python -m pytest tests/test_gpu_box.py -x -q 2>&1 | tail -5
