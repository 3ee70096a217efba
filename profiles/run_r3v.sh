cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err; echo bench rc=$?
grep "^\[bench" gpurun_out/r3v_bench.err | tail -5
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3v_ref.json 2> gpurun_out/r3v_ref.err; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-configs --no-scale > gpurun_out/r3v_plain.json 2>/dev/null || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3v_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-configs --no-scale > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:square_rs_kernel -s 5 -c 1 -o gpurun_out/r3v_rs -f python bench.py --steps 2 --warmup 1 --no-configs --no-scale > gpurun_out/r3v_ncu.log 2>&1
tail -2 gpurun_out/r3v_ncu.log
