timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest_all.log 2>&1; tail -4 gpurun_out/r2e_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -c 300 gpurun_out/r2e_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; tail -c 400 gpurun_out/r2e_bench_ref.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2e_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['steps'], d['warmup'])
for k in ('config3','config4','config5'):
    c=d.get(k,{})
    print(k, {kk:c.get(kk) for kk in ('ms','ms_outer','ms_fused_outer_vector','value','error')}, (c.get('roofline') or {}).get('frac'))
P
