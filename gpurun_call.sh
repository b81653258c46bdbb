TAILN=25 ./run_gpu_tests.sh fullsize
TAILN=30 ./run_gpu_tests.sh parity
b() { name=$1; shift; echo "=== bench $name"; DGMR_BENCH_DUMP=gpurun_out/shapes_$name.tsv timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; tail -c 1500 gpurun_out/bench_$name.json; tail -n 3 gpurun_out/bench_$name.err; }
b c3 --steps 5 --warmup 3
b c2 --config c2 --steps 10 --warmup 3 --no-cpu-baseline
b c2graph --config c2 --cuda-graph --steps 10 --warmup 3 --no-cpu-baseline
b dropin --mode dropin --steps 3 --warmup 1 --no-cpu-baseline
b c5 --config c5 --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu
b x3 --precision 3xtf32 --steps 3 --warmup 1 --no-cpu-baseline --no-ref-gpu
