mkdir -p gpurun_out/r2
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2/bench_n8_v2.json 2> gpurun_out/r2/bench_n8_v2.err)
python -c "
import json
for f in ('bench_n8_v2',):
    try:
        d=json.load(open('gpurun_out/r2/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('exchange'), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
"
tail -3 gpurun_out/r2/bench_n8_v2.err
