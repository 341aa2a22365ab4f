# round-2 final capture on one B200 (run through gpurun): parity tests, the bench line, the
# ncu launch list of the same bench command, one `--set full` capture of the four hb:: kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; tail -4 gpurun_out/r2_final_pytest.log
timeout 300 python bench.py > gpurun_out/r2_final_bench_n1.log 2>&1; tail -1 gpurun_out/r2_final_bench_n1.log | cut -c1-400
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-uniform > gpurun_out/r2_final_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'update_short_kernel|update_long_kernel|lookup_fwd_kernel|cluster_sort_runs_kernel' -s 12 -c 8 -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-uniform > gpurun_out/r2_final_ncu2.log 2>&1
ls -la gpurun_out | tail -6
