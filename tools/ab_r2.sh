# round-2 measurement helper (run under gpurun)
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2_tests7.log; tail -12 gpurun_out/r2_tests7.log
Q="--no-e2e --no-cpu --no-others --no-parity --steps 60"
run() { out=$(env JR_BENCH_CAPTURE=$2 $3 python bench.py $Q 2>gpurun_out/ab_$1_$2.err | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']/1e9,3), round(d['ms_per_step'],4), d['instructions_per_step'])" 2>&1 | tail -1); echo "$1 cap=$2 $3: $out" | tee -a gpurun_out/r2_ab.txt; }
rm -f gpurun_out/r2_ab.txt
run fold 1
run fold 0
run nofold 1 JR_NO_FOLD=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_fold.csv python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 6 --warmup 3 > gpurun_out/r2_ncu_fold.log 2>&1
python bench.py --steps 100 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -c 600 gpurun_out/r2_bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'e2e_dense',d['e2e_dense_input']['value'],'e2e_plain',d['e2e_no_output']['value'])
print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['other_configs'].items()}, d['variants'], [p['bit_exact'] for p in d['parity']])
print(d['e2e'])
"
