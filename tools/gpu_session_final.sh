#!/bin/bash
# last GPU call of the round: decode timing on the box + the CLI / sink tests (incl. the forced gather path)
mkdir -p gpurun_out/r2
make -s -C reconstruction_b200/host >/dev/null 2>&1
reconstruction_b200/host/reconstruction --decode-bench tmp_big.jpg 8 > gpurun_out/r2/final_decode.json 2>&1
nproc >> gpurun_out/r2/final_decode.json; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2/final_decode.json
timeout 150 python -m pytest tests/test_host_gpu.py tests/test_sink_gpu.py -m gpu -q -x --durations=8 > gpurun_out/r2/final_host_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2/final_host_tests.log
tail -5 gpurun_out/r2/final_host_tests.log
