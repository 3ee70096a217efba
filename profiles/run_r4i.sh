cd $GRAFT_REPO_ROOT
python profiles/r2_configs.py circ 2>&1 | cut -c1-110 | grep -v two_passes | grep land_sea
