#!/bin/bash
mkdir -p gpurun_out/r2
timeout 58 python tools/run_handover_ab.py gpurun_out/r2/handover_ab2.json 7 \
  reconstruction_b200/host/reconstruction:SB200_ALLGATHER=0 \
  reconstruction_b200/host/reconstruction_cur:SB200_ALLGATHER=0 \
  reconstruction_b200/host/reconstruction:SB200_ALLGATHER=0 > gpurun_out/r2/handover_ab2.log 2>&1
echo "rc=$?" >> gpurun_out/r2/handover_ab2.log
tail -c 300 gpurun_out/r2/handover_ab2.log
