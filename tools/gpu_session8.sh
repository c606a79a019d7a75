mkdir -p gpurun_out/r2
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2/bench_n8_v4_allgather.json 2> gpurun_out/r2/bench_n8_v4_allgather.err)
python -c "
import json
d=json.load(open('gpurun_out/r2/bench_n8_v4_allgather.json')); print('n8 allgather', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('exchange'))" || tail -5 gpurun_out/r2/bench_n8_v4_allgather.err
