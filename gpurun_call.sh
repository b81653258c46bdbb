TAILN=8 ./run_gpu_tests.sh allv smoke
b() { name=$1; shift; echo "=== bench $name"; timeout 1200 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d.get('reference_gpu_eager')); print(d.get('cpu_baseline')); print(d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'])"; tail -n 2 gpurun_out/bench_$name.err; }
b final2_default
