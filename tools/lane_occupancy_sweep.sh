for w in 2 4 8 12; do
  n=$((148*w*32))
  echo "== warps/SM $w reads $n"
  C3POA_LANE_WARPS_PER_SM=$w C3POA_GPU_LIB=c3poa_b200/libc3poa_gpu.so timeout 300 python tools/lane_prof_run.py $n 2>&1 | tail -1 | grep -o "'poa_ms': [0-9.]*"
done
