# round-2 measurement helper (run under gpurun)
Q="--no-e2e --no-cpu --no-others --no-parity --steps 60"
run() { out=$(env JR_BENCH_CAPTURE=$2 $3 python bench.py $Q 2>gpurun_out/ab_$1_$2.err | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']/1e9,3), round(d['ms_per_step'],4), d['instructions_per_step'])" 2>&1 | tail -1); echo "$1 cap=$2 $3: $out" | tee -a gpurun_out/r2_ab.txt; }
rm -f gpurun_out/r2_ab.txt
run fold 1
run fold 0
JR_BENCH_TRACE=1 python bench.py --no-cpu --no-others --no-parity --steps 100 > gpurun_out/r2_bench_e2e.json 2> gpurun_out/r2_bench_e2e.err; tail -c 300 gpurun_out/r2_bench_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e2e.json'))
print('value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'), e.get('records_per_step'), e.get('d2h_bytes_per_step'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_fold.csv python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 6 --warmup 3 > gpurun_out/r2_ncu_fold.log 2>&1
python -m pytest tests/test_sym_fold.py tests/test_stream_path.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2_tests8.log; tail -8 gpurun_out/r2_tests8.log
