timeout 300 python -m pytest tests/test_gpu_tc3.py tests/test_gpu_tc.py -q > gpurun_out/x25_tests.log 2>&1; tail -4 gpurun_out/x25_tests.log
timeout 200 python tools/time_fused.py > gpurun_out/x25_time.log 2>&1; tail -3 gpurun_out/x25_time.log
python tools/time_step.py > gpurun_out/x25_step.log 2>&1; tail -1 gpurun_out/x25_step.log
