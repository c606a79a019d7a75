mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r2/gputests_v2.log 2>&1
(timeout 200 python tools/time_stages.py 5 256 192 2 2>&1 | tail -32) > gpurun_out/r2/stages_v5.log 2>&1
(timeout 600 python bench.py > gpurun_out/r2/bench_n1_v2.json 2> gpurun_out/r2/bench_n1_v2.err)
tail -6 gpurun_out/r2/gputests_v2.log; grep -E "InitialMatch|DisparityRefine|total stages|refine sweeps" gpurun_out/r2/stages_v5.log; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_n1_v2.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_levels']['frac'], d['ncc_top_level'], d.get('parity'), d['one_pair_at_a_time']['ms_per_pair'], d['gpu_launches'], d['rectify'])"
tail -3 gpurun_out/r2/bench_n1_v2.err
