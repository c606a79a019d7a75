mkdir -p gpurun_out/r2
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r2/gputests_v3.log 2>&1
(timeout 400 python bench.py > gpurun_out/r2/bench_n1_v3.json 2> gpurun_out/r2/bench_n1_v3.err)
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v4.csv python tools/prof_pair.py 5 256 192 1 > gpurun_out/r2/launches_v4.log 2>&1
tail -3 gpurun_out/r2/gputests_v3.log; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_n1_v3.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['ncc_top_level']['ms_per_step'], d['ncc_top_level']['frac_of_hbm_peak'], d.get('parity',{}).get('px_differing'), d['one_pair_at_a_time']['ms_per_pair'], d['gpu_launches'])"
