cd $GRAFT_REPO_ROOT
python tools/_run36.py 2>&1 | tail -1 | cut -c1-250
C3POA_CONK_INT32=1 python tools/_run36.py 2>&1 | tail -1 | cut -c1-250
