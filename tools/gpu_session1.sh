set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k torture 2>&1 | tail -30
