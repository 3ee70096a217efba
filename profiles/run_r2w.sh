cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_engine.py tests/test_gpu_plugins.py tests/test_gpu_config2.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py square 2>&1 | cut -c1-110
