cd $GRAFT_REPO_ROOT
python bench.py --no-configs --steps 5 --warmup 3 > gpurun_out/r3o.json 2> gpurun_out/r3o.err; echo rc=$?
tail -25 gpurun_out/r3o.err
