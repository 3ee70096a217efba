cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_box.py -q -x -k two_pass 2>&1 | grep -E "^E|assert|passed|failed|Mismatch|Max abs|test_gpu_box.py:[0-9]+" | head -30
timeout 900 python -m pytest tests/test_gpu_config2.py tests/test_gpu_engine.py -q -x 2>&1 | tail -4
python profiles/r2_configs.py circ square 2>&1 | cut -c1-170 | grep -v two_passes
