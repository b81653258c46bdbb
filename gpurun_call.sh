TAILN=6 ./run_gpu_tests.sh parity
b() { name=$1; shift; echo "=== bench $name"; DGMR_BENCH_DUMP=gpurun_out/shapes_$name.tsv timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print({k:v for k,v in d['kernel_breakdown_ms'].items() if k in ('round_tf32','conv_umma')})"; tail -n 3 gpurun_out/bench_$name.err; }
b c3s --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline
