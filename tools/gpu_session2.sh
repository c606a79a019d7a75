mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_comm_gpu.py tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r2/n2_tests_v3.log 2>&1; tail -2 gpurun_out/r2/n2_tests_v3.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2/bench_n2_v2.json 2> gpurun_out/r2/bench_n2_v2.err); python -c "
import json
d=json.load(open('gpurun_out/r2/bench_n2_v2.json')); print('allgather', 'value', round(d['value']), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d.get('exchange'))" || tail -5 gpurun_out/r2/bench_n2_v2.err
