# round-2 measurement pass (run under gpurun): GPU tests, full bench, step breakdown, launch list, one full ncu capture of the hot kernel
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests.log 2>&1
tail -5 gpurun_out/r2_gpu_tests.log
( time python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err ) 2>&1 | tail -3
tail -c 600 gpurun_out/r2_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>&1; cut -c1-400 gpurun_out/r2_bench_reference.json
python tools/step_breakdown.py 60 > gpurun_out/r02_step_breakdown.txt 2>&1; cat gpurun_out/r02_step_breakdown.txt
Q="--no-e2e --no-cpu --no-others --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py $Q --steps 6 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sym2_kernel -s 12 -c 1 -f -o gpurun_out/r02d_sym2_kernel python bench.py $Q --steps 4 --warmup 3 > gpurun_out/r02d_ncu_sym2.log 2>&1
ls -la gpurun_out | head -30
