set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_prof_run.py 37888 > gpurun_out/r2_gprof1.txt 2>&1
cat gpurun_out/r2_gprof1.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:c3_poa_grp_kernel -c 1 -o gpurun_out/r2_grp_v1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu1.log
