mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/n8_gpus.txt; nproc >> gpurun_out/r2/n8_gpus.txt; free -g | head -2 >> gpurun_out/r2/n8_gpus.txt
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2/bench_n8_v1.json 2> gpurun_out/r2/bench_n8_v1.err)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 6 --warmup 3 --no-exchange > gpurun_out/r2/bench_n8_v1_noexchange.json 2> gpurun_out/r2/bench_n8_v1_noexchange.err)
(timeout 900 python tools/run_config_e.py gpurun_out/r2/configE_n8_v1.json 20 > gpurun_out/r2/configE_n8_v1.log 2>&1)
python -c "
import json
for f in ('bench_n8_v1','bench_n8_v1_noexchange'):
    try:
        d=json.load(open('gpurun_out/r2/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('exchange'), d['clocks'])
    except Exception as e: print(f, 'ERR', e)
"
tail -3 gpurun_out/r2/configE_n8_v1.log; tail -3 gpurun_out/r2/bench_n8_v1.err
