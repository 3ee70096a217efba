cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:square_|init_stats|collect_suspects|box_|edge_flags" -c 400 --csv --log-file gpurun_out/r4b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-configs --no-scale > /dev/null 2>&1
echo launches done; wc -l gpurun_out/r4b_launches_bench.csv
