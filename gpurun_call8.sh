mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "=== c3 x8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --no-ref-gpu --no-cpu-baseline > gpurun_out/bench8_c3b.json 2> gpurun_out/bench8_c3b.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench8_c3b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config'])"; tail -n 3 gpurun_out/bench8_c3b.err
