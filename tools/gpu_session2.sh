mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/n2_gpus.txt
(timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r2/n2_tests.log 2>&1
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2/bench_n2_v1.json 2> gpurun_out/r2/bench_n2_v1.err)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-exchange > gpurun_out/r2/bench_n2_v1_noexchange.json 2> gpurun_out/r2/bench_n2_v1_noexchange.err)
tail -5 gpurun_out/r2/n2_tests.log
python -c "
import json
for f in ('bench_n2_v1','bench_n2_v1_noexchange'):
    try:
        d=json.load(open('gpurun_out/r2/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('exchange'))
    except Exception as e: print(f, 'ERR', e)
"
tail -5 gpurun_out/r2/bench_n2_v1.err
