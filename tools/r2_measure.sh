# round-2 measurement pass (run under gpurun): GPU tests, full bench, launch list, one full ncu capture per hot kernel
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_gpu_tests.log
( time python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2_bench_full.err
python tools/step_breakdown.py 40 > gpurun_out/r2_step_breakdown.txt 2>&1; cat gpurun_out/r2_step_breakdown.txt
Q="--no-e2e --no-cpu --no-others --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py $Q --steps 6 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym_kernel -s 12 -c 1 -f -o gpurun_out/r02_sym_kernel python bench.py $Q --steps 4 --warmup 3 > gpurun_out/r02_ncu_sym.log 2>&1
JR_NO_FOLD=1 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -f -o gpurun_out/r02_step_kernel python bench.py $Q --steps 4 --warmup 3 > gpurun_out/r02_ncu_step.log 2>&1
ls -la gpurun_out
