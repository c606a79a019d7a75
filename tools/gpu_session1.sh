set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/time_stages.py 5 256 192 2 2>&1 | tail -22
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ncc|stats|ranges' --csv --log-file gpurun_out/launches_match.csv python tools/prof_pair.py 5 256 192 1 2>&1 | tail -1
