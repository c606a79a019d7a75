cd /root/repo
timeout 600 python tools/time_stages.py 5 256 192 3 2>&1 | grep -E " 9 DisparityRefine|RefineSweeps|match_pair"
