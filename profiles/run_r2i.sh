cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_box.py -x -q 2>&1 | tail -3
python profiles/sweep_box.py "" "IMPROVER_B200_LIB=$GRAFT_REPO_ROOT/improver_b200/lib/libimprover_b200_k8.so" > gpurun_out/r2i_sweep.jsonl 2>&1
cat gpurun_out/r2i_sweep.jsonl
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r2i_bench.err
head -c 1500 gpurun_out/r2i_bench.json
