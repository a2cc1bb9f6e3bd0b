cd $GRAFT_REPO_ROOT
for r in 0 4096 12288; do echo res_cap $r; C3POA_PK_RES=$r python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | grep -o "'peaks_ms': [0-9.]*"; done
