set -x
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_refine_fused -s 65 -c 1 -o gpurun_out/prof_fused_v8 python tools/prof_pair.py 5 256 192 1 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v8.csv python tools/prof_pair.py 5 256 192 2 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none -k regex:k_ncc_screen5 -s 6 -c 1 -o gpurun_out/prof_screen_v1 python tools/prof_pair.py 5 256 192 1 2>&1 | tail -1
