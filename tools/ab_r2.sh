mkdir -p gpurun_out
for i in 1 2; do
JR_BENCH_TRACE=1 python bench.py --no-cpu --no-others --no-parity --steps 150 > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; tail -c 300 gpurun_out/ab_bench.err
python -c "
import json; d=json.load(open('gpurun_out/ab_bench.json'))
print('value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'), e.get('host_fold_threads'))
"
done
