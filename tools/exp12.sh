python tools/trace_fused.py fwd > gpurun_out/x12_trace_fwd.log 2>&1; cat gpurun_out/x12_trace_fwd.log | head -12
python tools/trace_fused.py dgrad > gpurun_out/x12_trace_dgrad.log 2>&1; cat gpurun_out/x12_trace_dgrad.log | head -8
