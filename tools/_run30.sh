cd $GRAFT_REPO_ROOT
for v in lim6 lim7; do echo $v; C3POA_GPU_LIB=build/variants/lib_$v.so python tools/grp_ncu_run.py 100000 grp 2>&1 | tail -1 | grep -o "fast kernel.*"; done
