# round-2 measurement helper (run under gpurun)
Q="--no-cpu --no-others --no-parity --steps 100"
JR_BENCH_TRACE=1 python bench.py $Q > gpurun_out/r2_bench_e2e.json 2> gpurun_out/r2_bench_e2e.err; tail -c 300 gpurun_out/r2_bench_e2e.err
JR_BENCH_TRACE=1 JR_FOLD_THREADS=4 python bench.py $Q > gpurun_out/r2_bench_e2e_t4.json 2> gpurun_out/r2_bench_e2e_t4.err
for f in gpurun_out/r2_bench_e2e.json gpurun_out/r2_bench_e2e_t4.json; do python -c "
import json; d=json.load(open('$f'))
print('value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'), e.get('records_per_step'), e.get('d2h_bytes_per_step'), e.get('folded_groups_last_step'))
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_fold.csv python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 6 --warmup 3 > gpurun_out/r2_ncu_fold.log 2>&1
python -m pytest tests/test_sym_fold.py tests/test_stream_path.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2_tests10.log; tail -5 gpurun_out/r2_tests10.log
