cd /root/repo
for v in 0 2 7 1; do for T in 5 6; do echo "tile $v T $T"; SB200_REFINE_TILE=$v SB200_REFINE_T=$T timeout 600 python tools/time_stages.py 5 256 192 2 2>&1 | grep -E "RefineSweeps"; done; done
