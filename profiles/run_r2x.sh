cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py tests/test_gpu_engine.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py square 2>&1 | cut -c1-110 | grep r5
python profiles/run_box_once.py 5 mask > gpurun_out/r2x_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_box -c 1 -o gpurun_out/r2x_boxm -f python profiles/run_box_once.py 5 mask > gpurun_out/r2x_ncu.log 2>&1
tail -2 gpurun_out/r2x_ncu.log
