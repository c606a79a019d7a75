set -x
cd /root/repo
timeout 900 python tools/sweep_refine.py 5 256 192 "5:0,5:8,5:9,6:8" 2>&1 | tail -20
timeout 600 python tools/time_stages.py 5 256 192 2 2>&1 | tail -22
