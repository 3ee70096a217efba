cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_box.py tests/test_gpu_fullsize.py tests/test_gpu_lazy.py tests/test_gpu_plugins.py -q 2>&1 | tail -2
python profiles/run_multi_threshold.py 2>&1 | cut -c1-100
python bench.py --no-scale --steps 20 --warmup 3 > gpurun_out/r3z_bench.json 2>/dev/null; python -c "import json,sys; d=json.loads(open('gpurun_out/r3z_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['cube_ms'], d['burst']['ms_per_step']/8, d['max_abs'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']); print({k:round(v['ms'],4) for k,v in d['configs'].items()})"
