timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest_all.log 2>&1; tail -2 gpurun_out/r2g_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
for k in ('config3','config4','config5'):
    c=d.get(k,{})
    print(k, {kk:c.get(kk) for kk in ('ms','ms_outer','ms_fused_outer_vector','value','error')}, (c.get('roofline') or {}).get('frac'))
P
