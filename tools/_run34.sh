cd $GRAFT_REPO_ROOT
echo base; python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | grep -o "'peaks_ms': [0-9.]*"
for m in 5 6 8; do echo minb $m; C3POA_GPU_LIB=build/variants/lib_pk$m.so python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | grep -o "'peaks_ms': [0-9.]*"; done
