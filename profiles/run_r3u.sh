cd $GRAFT_REPO_ROOT
NBH_EXP_LS_DIRECT=1 python profiles/r2_configs.py circ 2>&1 | cut -c1-120 | grep -v two_passes | grep -E "r5_land|r9_land|r27_land"
