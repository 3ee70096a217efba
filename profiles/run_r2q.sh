cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_config2.py -q -x 2>&1 | tail -3
python profiles/run_box_once.py 27 > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2q_vh27_launches.csv python profiles/run_box_once.py 27 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2q_vh27m_launches.csv python profiles/run_box_once.py 27 mask > /dev/null 2>&1
python profiles/r2_configs.py circ 2>&1 | cut -c1-170 | grep -v two_passes
