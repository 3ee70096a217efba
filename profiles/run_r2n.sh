cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_box.py -q -x 2>&1 | tail -4
python profiles/run_circ_once.py 9 > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2n_circ9_launches.csv python profiles/run_circ_once.py 9 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2n_circ81_launches.csv python profiles/run_circ_once.py 81 48 > /dev/null 2>&1
python profiles/r2_configs.py square 2>&1 | cut -c1-170
