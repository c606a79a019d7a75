set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in 8 9 10 11; do SB200_REFINE_TILE=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "live_oracle or torture" 2>&1 | tail -1; done
timeout 900 python tools/sweep_refine.py 5 256 192 "5:1,5:8,5:9,5:10,5:11" 2>&1 | tail -6
