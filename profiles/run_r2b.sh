mkdir -p gpurun_out
for pb in 4096 8192 16384 32768; do
HB_PIECE_BYTES=$pb timeout 120 python bench.py --steps 30 --no-cpu-baseline --no-e2e --no-uniform 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('piece_bytes',$pb, l['ms_per_step'], {k:round(v['ms_avg']*1e3,1) for k,v in l['kernels'].items()})"
done 2>&1 | tee gpurun_out/r2_piece_sweep.log
