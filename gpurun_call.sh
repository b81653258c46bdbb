echo "=== upconv tests"; timeout 300 python -m pytest tests/test_umma_gpu.py -x -q -m gpu -k "upconv or subpix or pairconv" --timeout 120 > gpurun_out/upconv.log 2>&1; echo "exit $?"; tail -12 gpurun_out/upconv.log
timeout 200 python tests/time_subpix_rows.py 2>&1 | tail -4
