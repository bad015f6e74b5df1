mkdir -p gpurun_out
run() {
python bench.py --no-cpu --no-others --no-parity --steps 100 > gpurun_out/ab_bench_$1.json 2> gpurun_out/ab_bench.err; tail -c 300 gpurun_out/ab_bench.err
python -c "
import json; d=json.load(open('gpurun_out/ab_bench_$1.json'))
print('$1 value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'), e.get('host_fold_threads'))
"
}
export JR_BENCH_TRACE=1
run two
JR_E2E_ONE_THREAD=1 run one
JR_FOLD_THREADS=4 run two_t4
