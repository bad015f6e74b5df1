# round-2 measurement helper (run under gpurun): quick device-resident A/B + tests + full bench
Q="--no-e2e --no-cpu --no-others --no-parity --steps 60"
run() { # name lib cap extra-env
  out=$(env JR_ENGINE_LIB=$2 JR_BENCH_CAPTURE=$3 $4 python bench.py $Q 2>gpurun_out/ab_$1_$3.err | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']/1e9,3), round(d['ms_per_step'],4), d['instructions_per_step'])" 2>&1 | tail -1)
  echo "$1 cap=$3 $4: $out" | tee -a gpurun_out/r2_ab.txt
}
rm -f gpurun_out/r2_ab.txt
L=josefine_b200/csrc
run fold $L/libjosefine_b200.so 1
run fold $L/libjosefine_b200.so 0
run nofold $L/libjosefine_b200.so 1 JR_NO_FOLD=1
run nofold $L/libjosefine_b200.so 0 JR_NO_FOLD=1
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_tests5.log; tail -4 gpurun_out/r2_tests5.log
python bench.py --steps 100 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -c 600 gpurun_out/r2_bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_full.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'e2e_plain',d['e2e_no_output']['value'], 'cpu', d['cpu_baseline'])
print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d['other_configs'].items()}, d['variants'], [p['bit_exact'] for p in d['parity']])
print(d['roofline'])
"
