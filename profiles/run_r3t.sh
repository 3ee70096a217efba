cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_config2.py -q -k "27 or 9" 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py circ 2>&1 | cut -c1-120 | grep -v two_passes | grep -E "r27|r9_"
