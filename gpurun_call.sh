TAILN=25 ./run_gpu_tests.sh allv
grep -E "DISC" gpurun_out/allv.log | head -20
