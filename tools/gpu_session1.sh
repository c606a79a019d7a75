set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/sweep_refine.py 5 256 192 "5:-1,5:1,5:0" 2>&1 | tail -4
SB200_REFINE_TMA=0 timeout 900 python tools/sweep_refine.py 5 256 192 "5:-1" 2>&1 | tail -1
