echo "=== row wgrad tests"; timeout 600 python -m pytest tests/test_umma_gpu.py -x -q -m gpu -k "wgrad or upconv or subpix" --timeout 300 > gpurun_out/wgrad.log 2>&1; echo "exit $?"; tail -15 gpurun_out/wgrad.log
echo "=== time_wgrad_narrow"; timeout 300 python tests/time_wgrad_narrow.py > gpurun_out/time_wgrad_narrow.log 2>&1; echo "exit $?"; cat gpurun_out/time_wgrad_narrow.log
