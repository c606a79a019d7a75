set -x
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_refine_fused -s 65 -c 1 -o gpurun_out/prof_fused_v9 python tools/prof_pair.py 5 256 192 1 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v9.csv python tools/prof_pair.py 5 256 192 2 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()"
