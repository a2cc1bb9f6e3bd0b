cd $GRAFT_REPO_ROOT
C3POA_GRP_TIMING=1 python tools/_run36.py 2>&1 | grep "poa stage\|host:" | paste - - | cut -c1-150 | tail -30
