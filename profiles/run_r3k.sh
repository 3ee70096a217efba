cd $GRAFT_REPO_ROOT
python profiles/run_box_once.py 27 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_[vh]_kernel -c 2 -o gpurun_out/r3k_vh -f python profiles/run_box_once.py 27 > gpurun_out/r3k_ncu.log 2>&1
tail -2 gpurun_out/r3k_ncu.log
