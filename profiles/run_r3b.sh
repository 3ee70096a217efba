cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_config2.py tests/test_gpu_engine.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py square 2>&1 | cut -c1-110 | grep r27
python profiles/run_multi_threshold.py 2>&1 | cut -c1-120
python profiles/run_box_once.py 27 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r3b_vh27_launches.csv python profiles/run_box_once.py 27 > /dev/null 2>&1
