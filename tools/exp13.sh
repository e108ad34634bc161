timeout 300 python -m pytest tests/test_gpu_tc3.py -q > gpurun_out/x13_tests.log 2>&1; tail -5 gpurun_out/x13_tests.log
timeout 200 python tools/time_fused.py > gpurun_out/x13_time.log 2>&1; tail -4 gpurun_out/x13_time.log
python tools/trace_fused.py fwd > gpurun_out/x13_trace_fwd.log 2>&1; cat gpurun_out/x13_trace_fwd.log | head -4
python tools/trace_fused.py dgrad > gpurun_out/x13_trace_dgrad.log 2>&1; cat gpurun_out/x13_trace_dgrad.log | head -3
