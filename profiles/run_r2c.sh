set -x
cd $GRAFT_REPO_ROOT
python profiles/run_box_once.py 5 > gpurun_out/r2c_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_box -c 1 -o gpurun_out/r2c_box -f python profiles/run_box_once.py 5 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
