cd $GRAFT_REPO_ROOT
python profiles/sweep_box.py "" "IMPROVER_B200_LIB=$GRAFT_REPO_ROOT/improver_b200/lib/libimprover_b200_fake.so" 2>&1 | cut -c1-200
python profiles/dbg_e2e.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_box.py -x -q 2>&1 | tail -4
python profiles/run_multi_threshold.py 2>&1 | tail -5
