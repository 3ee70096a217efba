cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo bench rc=$?
tail -c 600 gpurun_out/r3e_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3e_ref.json 2> gpurun_out/r3e_ref.err; echo ref rc=$?
cut -c1-600 gpurun_out/r3e_ref.json
