cd $GRAFT_REPO_ROOT
for cap in 13 26 39 80; do
echo cap $cap
NBH_EXP_CAP=$cap python bench.py --no-configs --no-scale --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['cube_ms'], d['burst']['ms_per_step']/8, d['max_abs'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done
