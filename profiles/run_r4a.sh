cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; echo bench rc=$?
grep "^\[bench" gpurun_out/r4a_bench.err | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4a_ref.json 2> gpurun_out/r4a_ref.err; echo ref rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nbh -c 400 --csv --log-file gpurun_out/r4a_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-configs --no-scale > /dev/null 2>&1
echo launches done
