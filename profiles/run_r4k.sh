cd $GRAFT_REPO_ROOT
for w in 1000 300 3000 100; do
NBH_EXP_WAIT=$w python bench.py --no-configs --no-scale --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($w, d['cube_ms'], d['burst']['ms_per_step']/8, d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done
