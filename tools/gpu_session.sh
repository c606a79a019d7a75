mkdir -p gpurun_out/r2
(timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_comm_gpu.py "tests/test_gpu_baseline_parity.py::test_every_dump_point_at_size" -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2/parity4.log 2>&1
(timeout 200 python tools/time_stages.py 5 256 192 2 2>&1 | tail -32) > gpurun_out/r2/stages_v3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v3.csv python tools/prof_pair.py 5 256 192 1 > gpurun_out/r2/launches_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ncc_band -s 6 -c 1 -o gpurun_out/r2/ncc_band_v3 -f python tools/prof_pair.py 5 256 192 1 > gpurun_out/r2/ncu_band_v3.log 2>&1
tail -5 gpurun_out/r2/parity4.log; grep -E "InitialMatch|total stages" gpurun_out/r2/stages_v3.log
