cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_box.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py square 2>&1 | cut -c1-110 | grep r5
python profiles/run_box_once.py 5 > gpurun_out/r2v_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_box -c 1 -o gpurun_out/r2v_box -f python profiles/run_box_once.py 5 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
