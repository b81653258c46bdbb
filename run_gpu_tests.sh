#!/bin/bash
# GPU-box test driver: each stage in its own process (a CUDA fault in one cannot poison the next), bounded by `timeout`.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n "${TAILN:-15}" gpurun_out/$name.log; }
for stage in "$@"; do
  case $stage in
    kernels)  run kernels python -m pytest tests/test_kernels_gpu.py -x -q -m gpu --timeout 300 ;;
    umma)     run umma python -m pytest tests/test_umma_gpu.py -q -m gpu --timeout 300 ;;
    parity)   run parity python -m pytest tests/test_parity_gpu.py -q -s -m gpu --timeout 900 ;;
    fullsize) run fullsize python -m pytest tests/test_fullsize_gpu.py -q -s -m gpu --timeout 600 ;;
    all)      run all python -m pytest tests -x -q -m gpu --timeout 900 ;;
    allv)     run allv python -m pytest tests -q -m gpu --timeout 900 ;;
    smoke)    run smoke python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench)    run bench python bench.py --steps 3 --warmup 3 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
