TAILN=12 ./run_gpu_tests.sh kernels umma
TAILN=20 ./run_gpu_tests.sh fullsize
TAILN=30 ./run_gpu_tests.sh parity
echo "=== gru timing"; timeout 300 python tests/time_gru_conv.py 2>&1 | tee gpurun_out/time_gru.log | tail -12
b() { name=$1; shift; echo "=== bench $name"; DGMR_BENCH_DUMP=gpurun_out/shapes_$name.tsv timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['kernel_breakdown_ms'])"; tail -n 3 gpurun_out/bench_$name.err; }
b c3b --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline
b c2b --config c2 --steps 10 --warmup 3 --no-cpu-baseline
b c2graphb --config c2 --cuda-graph --steps 10 --warmup 3 --no-cpu-baseline
