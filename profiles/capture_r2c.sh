mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'cluster_sort_runs_kernel' -s 4 -c 2 -o gpurun_out/prof_r2c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out | tail -3
