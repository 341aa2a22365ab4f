# round-2 capture: calibration, fwd sweep (TMA vs register path), ncu full set of the hot kernels
mkdir -p gpurun_out
./tools/gather_peak > gpurun_out/r2_gather_peak.log 2>&1; tail -20 gpurun_out/r2_gather_peak.log
for d in 32 64 128; do for t in 0 1; do
  echo "dim=$d HB_NO_TMA=$t" >> gpurun_out/r2_fwd_sweep.log
  HB_NO_TMA=$t timeout 120 python bench.py --mode fwd --dim $d --max-rows 20000000 --steps 30 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], {k:v['ms_avg'] for k,v in l['kernels'].items()})" >> gpurun_out/r2_fwd_sweep.log 2>&1
done; done
cat gpurun_out/r2_fwd_sweep.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'update_apply_kernel|lookup_rows_tma_kernel|bucket_pass_kernel|bucket_hist_kernel|runs_kernel|queue_kernel' -s 56 -c 14 -o gpurun_out/prof_r2a python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_b_ncu2.log 2>&1
ls -la gpurun_out | tail -12
