mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2/parity2.log 2>&1
(timeout 60 compute-sanitizer --tool memcheck --print-limit 8 python tools/prof_pair.py 3 64 48 1 2>&1 | grep -E "SUMMARY|Error|points|at |by " | head -30) > gpurun_out/r2/memcheck2.log 2>&1
(timeout 200 python tools/time_stages.py 5 256 192 2 2>&1 | tail -40) > gpurun_out/r2/stages_band1.log 2>&1
(SB200_BAND=0 timeout 200 python tools/time_stages.py 5 256 192 2 2>&1 | tail -40) > gpurun_out/r2/stages_band0.log 2>&1
tail -8 gpurun_out/r2/parity2.log; cat gpurun_out/r2/memcheck2.log | tail -8; grep -E "InitialMatch|total stages" gpurun_out/r2/stages_band1.log gpurun_out/r2/stages_band0.log
