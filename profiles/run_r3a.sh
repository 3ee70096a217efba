cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_box.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/run_multi_threshold.py 2>&1 | cut -c1-120
