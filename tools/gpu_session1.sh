set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_stages.py 5 256 192 3 2>&1 | grep -E "Uniqueness|match_pair|level totals"
