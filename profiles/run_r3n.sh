cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python bench.py --no-configs --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['scale'])[:900]); print(d['value'], d['roofline']['frac'])"
