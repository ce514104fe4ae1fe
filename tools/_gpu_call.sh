timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; tail -c 300 gpurun_out/r2g_bench_n8.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_bench_n8.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'])
print('strong', d.get('strong',{}).get('ms_per_step'), d.get('strong',{}).get('value'))
for k in ('config3','config4','config5'):
    c=d.get(k,{})
    print(k, {kk:c.get(kk) for kk in ('ms','ms_outer','ms_fused_outer_vector','value','error','balance')}, (c.get('roofline') or {}).get('frac'))
P
