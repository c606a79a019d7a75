set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu.log 2>&1; tail -2 gpurun_out/bench_ncu.log
