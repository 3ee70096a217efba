cd $GRAFT_REPO_ROOT
for cap in 16 24 44; do
echo mtcap $cap
NBH_EXP_MTCAP=$cap python profiles/run_multi_threshold.py 2>&1 | cut -c1-100
done
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_box.py -q 2>&1 | tail -2
