mkdir -p gpurun_out
run8() { name=$1; shift; echo "=== $name"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" > gpurun_out/bench8_$name.json 2> gpurun_out/bench8_$name.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench8_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config'])"; tail -n 3 gpurun_out/bench8_$name.err; }
run8 c5 --config c5 --steps 3 --warmup 3
run8 c3 --steps 10 --warmup 3
