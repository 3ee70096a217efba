cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python profiles/r2_configs.py square 2>&1 | cut -c1-170
python profiles/run_box_once.py 27 > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2r_vh27_launches.csv python profiles/run_box_once.py 27 > /dev/null 2>&1
python profiles/r2_configs.py circ 2>&1 | cut -c1-170 | grep -v two_passes
