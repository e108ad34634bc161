export ABN_LIB=$PWD/abnet3_b200/libabnet3_b200_dbg.so
python tools/time_chain.py 0 1 2 4 6 8 9 > gpurun_out/x1_chain.log 2>&1
unset ABN_LIB
python tools/trace_tc2.py chain_nodep > gpurun_out/x1_trace_nodep.log 2>&1
python tools/trace_tc2.py chain > gpurun_out/x1_trace_chain.log 2>&1
cat gpurun_out/x1_chain.log
