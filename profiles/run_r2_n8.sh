mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2g_bench_n8.log 2>&1; tail -1 gpurun_out/r2g_bench_n8.log | python -c "
import sys,json; l=json.loads(sys.stdin.read())
print(l['value'], l['ms_per_step'])
for k,v in l['kernels'].items(): print(k, v['launches']/30, round(v['ms_avg']*1e3,1), round(v['ms_total']/30*1e3,1))
print(l['extra']); print([(r['kernel'], round(r['frac'],3)) for r in l['roofline_all']]); print(l['e2e'])" || tail -30 gpurun_out/r2g_bench_n8.log
