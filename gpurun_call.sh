echo "=== kwstack tests"; timeout 300 python -m pytest tests/test_umma_gpu.py -x -q -m gpu -k kwstack --timeout 120 > gpurun_out/kwstack.log 2>&1; echo "exit $?"; tail -15 gpurun_out/kwstack.log
echo "=== time_kwstack"; timeout 120 python tests/time_kwstack.py > gpurun_out/time_kwstack.log 2>&1; echo "exit $?"; cat gpurun_out/time_kwstack.log
echo "=== kernels"; timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu --timeout 120 > gpurun_out/kernels.log 2>&1; echo "exit $?"; tail -3 gpurun_out/kernels.log
b() { name=$1; shift; echo "=== bench $name"; DGMR_BENCH_DUMP=gpurun_out/shapes_$name.tsv timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['kernel_breakdown_ms'])"; tail -n 3 gpurun_out/bench_$name.err; }
b c3h --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline
