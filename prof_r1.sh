mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sparse_update_kernel|lookup_fwd_kernel|bucket_pass_kernel|bucket_hist_kernel|sparse_update_fixup' -s 30 -c 8 -o gpurun_out/prof_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
