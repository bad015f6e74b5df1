# round-2 measurement helper (run under gpurun)
python -m pytest tests/test_sym_fold.py tests/test_stream_path.py -m gpu -q -x --tb=short 2>&1 | tail -40 > gpurun_out/r2_tests6.log; tail -30 gpurun_out/r2_tests6.log
Q="--no-e2e --no-cpu --no-others --no-parity --steps 6 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_fold.csv python bench.py $Q > gpurun_out/r2_ncu_fold.log 2>&1
JR_NO_FOLD=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_nofold.csv python bench.py $Q > gpurun_out/r2_ncu_nofold.log 2>&1
