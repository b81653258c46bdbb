TAILN=8 ./run_gpu_tests.sh allv smoke
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3600 -c 3400 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 1 --no-ref-gpu --no-cpu-baseline > gpurun_out/launches_r02b.log 2>&1; echo "exit $?"; wc -l gpurun_out/launches_r02b.csv
b() { name=$1; shift; echo "=== bench $name"; timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['traffic'], d['roofline']['frac'])"; tail -n 3 gpurun_out/bench_$name.err; }
b c3q --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline
