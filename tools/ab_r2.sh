# round-2 A/B helper (run under gpurun)
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max; nproc
run() {
python bench.py --no-cpu --no-others --no-parity --steps 100 > gpurun_out/ab_bench_$1.json 2> gpurun_out/ab_bench.err; tail -c 300 gpurun_out/ab_bench.err
python -c "
import json; d=json.load(open('gpurun_out/ab_bench_$1.json'))
print('$1 value',d['value'],d['ms_per_step'])
for k in ('e2e','e2e_dense_input','e2e_no_output'):
    e=d[k]; print(k, e['value'], e['ms_per_step'], e.get('host_ms_per_step'))
"
}
export JR_BENCH_TRACE=1
JR_FOLD_THREADS=4 run t4
JR_FOLD_THREADS=8 run t8
JR_FOLD_THREADS=8 JR_FOLD_PIN=1 run t8pin
JR_FOLD_THREADS=4 JR_FOLD_PIN=1 run t4pin
JR_FOLD_THREADS=2 run t2
