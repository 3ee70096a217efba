cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_box.py -q 2>&1 | grep -E "^E|passed|failed" | head -20
python profiles/r2_configs.py square 2>&1 | cut -c1-110 | grep r27
python profiles/run_box_once.py 27 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_[vh]_kernel -c 2 -o gpurun_out/r3d_vh -f python profiles/run_box_once.py 27 > gpurun_out/r3d_ncu.log 2>&1
tail -2 gpurun_out/r3d_ncu.log
