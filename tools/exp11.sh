timeout 200 python -m pytest tests/test_gpu_tc3.py -q -k "forward" > gpurun_out/x11_fwd.log 2>&1; tail -12 gpurun_out/x11_fwd.log
timeout 200 python -m pytest tests/test_gpu_tc3.py -q -k "dgrad" > gpurun_out/x11_dgrad.log 2>&1; tail -12 gpurun_out/x11_dgrad.log
timeout 200 python tools/time_fused.py > gpurun_out/x11_time.log 2>&1; tail -4 gpurun_out/x11_time.log
