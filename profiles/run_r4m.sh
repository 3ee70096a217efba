cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -q -k "percentile or Percentile" 2>&1 | tail -1
