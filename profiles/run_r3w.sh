cd $GRAFT_REPO_ROOT
python profiles/run_multi_threshold.py 2>&1 | cut -c1-100 | head -2
NBH_EXP_MT1=1 python profiles/run_multi_threshold.py 2>&1 | cut -c1-100 | head -1
NBH_EXP_MT1=1 python bench.py --no-configs --no-scale --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['cube_ms'], d['max_abs'], json.dumps(d['roofline'])[:300])"
