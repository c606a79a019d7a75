mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_host_gpu.py tests/test_sink_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2/n2_tests_v2.log 2>&1; tail -3 gpurun_out/r2/n2_tests_v2.log
run() { tag=$1; shift; (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 8 --warmup 3 "$@" > gpurun_out/r2/dbg6_$tag.json 2> gpurun_out/r2/dbg6_$tag.err); python -c "
import json,sys
d=json.load(open('gpurun_out/r2/dbg6_$tag.json')); print('$tag', 'value', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), d.get('exchange'), d.get('sink_filter'))"; }
run sendrecv
SB200_NCCL_MAX_CTAS=16 run sendrecv16
