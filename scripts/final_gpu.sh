set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest6.log 2>&1; tail -2 gpurun_out/r02_gputest6.log
python bench.py > gpurun_out/r02_bench_final4.json 2> gpurun_out/r02_bench_final4.err; echo bench rc=$?
ncu --set full --clock-control none --import-source on -k regex:fused_tile_kernel -s 4 -c 1 -f -o gpurun_out/r02_fused_iid python scripts/run_fused.py 192x640 12 T 6 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_tile_kernel -s 4 -c 1 -f -o gpurun_out/r02_fused_375x1242 python scripts/run_fused.py 375x1242 12 T 6 > gpurun_out/ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launch_list.csv python bench.py --steps 8 --warmup 3 --no-train-step --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
tail -n 2 gpurun_out/ncu_a.log; tail -n 2 gpurun_out/ncu_b.log
