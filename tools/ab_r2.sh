Q="--no-e2e --no-cpu --no-others --no-parity --steps 60"
run() { # name lib cap extra-env
  out=$(env JR_ENGINE_LIB=$2 JR_BENCH_CAPTURE=$3 $4 python bench.py $Q 2>gpurun_out/ab_$1.err | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']/1e9,3), round(d['ms_per_step'],4))" 2>&1 | tail -1)
  echo "$1 cap=$3 $4: $out" | tee -a gpurun_out/r2_ab.txt
}
rm -f gpurun_out/r2_ab.txt
L=josefine_b200/csrc
run main $L/libjosefine_b200.so 0
run main $L/libjosefine_b200.so 1
run v1_noenc $L/ab/lib_v1.so 0
run v3_noinl $L/ab/lib_v3.so 0
run v3_noinl $L/ab/lib_v3.so 1
run v4_3cta $L/ab/lib_v4.so 0
run v4_3cta $L/ab/lib_v4.so 1
run main_W2 $L/libjosefine_b200.so 1 JR_TABLE_CACHE=2
run main_Us5 $L/libjosefine_b200.so 1 JR_SMEM_UNITS=5
run v4_W2 $L/ab/lib_v4.so 1 JR_TABLE_CACHE=2
python tools/phase_profile.py > gpurun_out/r2_phase_cap.txt 2>&1
JR_BENCH_CAPTURE=0 python tools/phase_profile.py > gpurun_out/r2_phase_nocap.txt 2>&1
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_tests3.log; tail -3 gpurun_out/r2_tests3.log
