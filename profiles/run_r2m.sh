cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_config2.py -q -x 2>&1 | tail -3
python profiles/dbg_circ2.py 2>&1 | tail -4
python profiles/r2_configs.py circ > gpurun_out/r2m_configs.jsonl 2>&1; cut -c1-160 gpurun_out/r2m_configs.jsonl
