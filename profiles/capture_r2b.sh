mkdir -p gpurun_out
HB_SHORT_OCC=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'update_short_kernel|update_long_kernel|lookup_fwd_kernel|runs_kernel' -s 16 -c 8 -o gpurun_out/prof_r2b python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out | tail -5
