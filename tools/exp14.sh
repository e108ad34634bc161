python tools/trace_fused.py fwd > gpurun_out/x14_trace_fwd.log 2>&1; cut -c 1-20,150- gpurun_out/x14_trace_fwd.log | head -4
python tools/trace_fused.py dgrad > gpurun_out/x14_trace_dgrad.log 2>&1; cut -c 1-20,150- gpurun_out/x14_trace_dgrad.log | head -3
