mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench_final_n1.log 2>&1; tail -1 gpurun_out/bench_final_n1.log | cut -c1-200
timeout 120 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_final_ref.log 2>&1; tail -1 gpurun_out/bench_final_ref.log | cut -c1-200
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu1.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:'sparse_update_kernel|lookup_fwd_kernel|bucket_pass_kernel|bucket_hist_kernel|sparse_update_fixup' -s 40 -c 8 -o gpurun_out/prof_r1c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2.log 2>&1
ls gpurun_out | tr '\n' ' '
