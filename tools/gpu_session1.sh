set -x
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_stages.py 5 256 192 3 2>&1 | grep -E " 2 InitialMatch| 6 Rematch|match_pair|evals"
