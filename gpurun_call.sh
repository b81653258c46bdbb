b() { name=$1; shift; echo "=== bench $name"; timeout 1200 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print({k:d.get(k) for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','reference_gpu_eager','impl','config')})"; tail -n 2 gpurun_out/bench_$name.err; }
b final_default
b final_reference --impl reference --steps 2 --warmup 1
b final_c2graph --config c2 --cuda-graph --steps 20 --warmup 5 --no-ref-gpu --no-cpu-baseline
b final_c2 --config c2 --steps 20 --warmup 5 --no-ref-gpu --no-cpu-baseline
b final_c5 --config c5 --steps 3 --warmup 3 --no-ref-gpu --no-cpu-baseline
b final_dropin --mode dropin --steps 3 --warmup 3 --no-ref-gpu --no-cpu-baseline
b final_x3 --precision 3xtf32 --steps 3 --warmup 3 --no-ref-gpu --no-cpu-baseline
