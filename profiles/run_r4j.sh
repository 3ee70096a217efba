cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_config2.py -q -k "9 or 45" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_box.py -q 2>&1 | tail -1
