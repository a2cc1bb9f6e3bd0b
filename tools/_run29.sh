cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b29.log 2>&1
tail -c 300 gpurun_out/r2_b29.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:c3_poa_grp_dp -s 1 -c 1 -o gpurun_out/r02_dp_full -f python tools/grp_ncu_run.py 37888 grp > gpurun_out/r2_run29a.txt 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:c3_poa_graph_kernel -s 1 -c 1 -o gpurun_out/r02_graph_full -f python tools/grp_ncu_run.py 37888 grp > gpurun_out/r2_run29b.txt 2>&1
tail -n 2 gpurun_out/r2_run29a.txt; tail -n 2 gpurun_out/r2_run29b.txt
