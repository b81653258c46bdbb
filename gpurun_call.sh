TAILN=12 ./run_gpu_tests.sh kernels umma fullsize
TAILN=30 ./run_gpu_tests.sh parity
echo "=== flags timing"; timeout 300 python tests/time_flags.py 2>&1 | tee gpurun_out/time_flags.log | tail -8
b() { name=$1; shift; echo "=== bench $name"; DGMR_BENCH_DUMP=gpurun_out/shapes_$name.tsv timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['kernel_breakdown_ms'])"; tail -n 3 gpurun_out/bench_$name.err; }
b c3d --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_umma|conv_subpix' -s 8 -c 8 -o gpurun_out/prof_r02 -f python tests/prof_kernels.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/ncu_full.log
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-ref-gpu --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "exit $?"; wc -l gpurun_out/launches_r02.csv
