cd $GRAFT_REPO_ROOT
python profiles/dbg_circ2.py 2>&1 | tail -6
