mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8 ) 2>&1 | tail -12
echo "== fused"; timeout 300 python tools/step_breakdown.py 40 --only 3 2>&1 | tee gpurun_out/ab_fused.txt
echo "== explicit truncate"; timeout 300 python tools/step_breakdown.py 40 --only 1 --explicit-truncate 2>&1 | tee gpurun_out/ab_explicit.txt
JR_BENCH_TRACE=1 python bench.py --no-cpu --no-others --no-parity --steps 100 > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; tail -c 300 gpurun_out/ab_bench.err
python -c "
import json; d=json.load(open('gpurun_out/ab_bench.json'))
print('value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'))
"
