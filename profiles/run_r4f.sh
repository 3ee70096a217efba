cd $GRAFT_REPO_ROOT
python profiles/run_circ_once.py 9 > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:circ_gather4_kernel|circ_prefix_kernel" -s 2 -c 2 -o gpurun_out/r4f_circ -f python profiles/run_circ_once.py 9 > gpurun_out/r4f_ncu1.log 2>&1
tail -1 gpurun_out/r4f_ncu1.log
python profiles/run_pct_once.py > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:percentile_kernel -c 1 -o gpurun_out/r4f_pct -f python profiles/run_pct_once.py > gpurun_out/r4f_ncu2.log 2>&1
tail -1 gpurun_out/r4f_ncu2.log
