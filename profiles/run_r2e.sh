cd $GRAFT_REPO_ROOT
python profiles/sweep_box.py "" "IMPROVER_B200_LIB=$GRAFT_REPO_ROOT/improver_b200/lib/libimprover_b200_b2.so" "NBH_BOX_STAGES=2" > gpurun_out/r2e_sweep.jsonl 2>&1
cat gpurun_out/r2e_sweep.jsonl
timeout 300 python -m pytest tests/test_gpu_box.py -x -q 2>&1 | tail -5
IMPROVER_B200_LIB=$GRAFT_REPO_ROOT/improver_b200/lib/libimprover_b200_b2.so timeout 300 python -m pytest tests/test_gpu_box.py -x -q 2>&1 | tail -3
