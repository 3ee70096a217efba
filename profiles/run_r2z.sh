cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_box.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/run_multi_threshold.py 2>&1 | cut -c1-120
python profiles/run_mt_once.py 4 > gpurun_out/r2z_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_mt -c 1 -o gpurun_out/r2z_mt -f python profiles/run_mt_once.py 4 > gpurun_out/r2z_ncu.log 2>&1
tail -2 gpurun_out/r2z_ncu.log
python profiles/run_box_once.py 27 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2z_vh27_launches.csv python profiles/run_box_once.py 27 > /dev/null 2>&1
