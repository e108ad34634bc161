export PADTO=64
python tools/trace_tc2.py chain > gpurun_out/x4_trace_chain.log 2>&1
ABN_GEMM_GRID=128 python tools/trace_tc2.py chain > gpurun_out/x4_trace_chain128.log 2>&1
python tools/trace_tc2.py chain_nodep > gpurun_out/x4_trace_nodep.log 2>&1
