# round-2 measurement helper (run under gpurun): quick device-resident A/B + launch list + one full ncu capture
Q="--no-e2e --no-cpu --no-others --no-parity --steps 60"
run() { # name lib cap extra-env
  out=$(env JR_ENGINE_LIB=$2 JR_BENCH_CAPTURE=$3 $4 python bench.py $Q 2>gpurun_out/ab_$1_$3.err | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']/1e9,3), round(d['ms_per_step'],4), d['instructions_per_step'])" 2>&1 | tail -1)
  echo "$1 cap=$3 $4: $out" | tee -a gpurun_out/r2_ab.txt
}
rm -f gpurun_out/r2_ab.txt
L=josefine_b200/csrc
run main $L/libjosefine_b200.so 0
run main $L/libjosefine_b200.so 1
run main_noparts $L/libjosefine_b200.so 1 JR_PARTS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 6 --warmup 3 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/r2_step --force-overwrite python bench.py --no-e2e --no-cpu --no-others --no-parity --steps 4 --warmup 3 > gpurun_out/r2_ncu_full.log 2>&1
python tools/phase_profile.py > gpurun_out/r2_phase_cap.txt 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_tests4.log; tail -3 gpurun_out/r2_tests4.log
