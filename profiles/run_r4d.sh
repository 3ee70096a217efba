cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_config2.py tests/test_gpu_edge_cases.py -q -x 2>&1 | tail -2
python profiles/r2_configs.py circ 2>&1 | cut -c1-110 | grep -v two_passes | grep -E "r9_|r5_"
