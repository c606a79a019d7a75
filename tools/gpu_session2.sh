mkdir -p gpurun_out/r2
run() { tag=$1; shift; (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 8 --warmup 3 "$@" > gpurun_out/r2/dbg5_$tag.json 2> gpurun_out/r2/dbg5_$tag.err); python -c "
import json,sys
d=json.load(open('gpurun_out/r2/dbg5_$tag.json')); print('$tag', 'value', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), d.get('exchange'))"; }
SB200_BENCH_SLOTS=2 run slots2
SB200_BENCH_SLOTS=3 run slots3
SB200_BENCH_SLOTS=5 run slots5
run noex --no-exchange
