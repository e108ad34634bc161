for p in 8 64; do echo "PADTO=$p"; PADTO=$p python tools/time_chain.py 0; done > gpurun_out/x2_pad.log 2>&1
export ABN_LIB=$PWD/abnet3_b200/libabnet3_b200_dbg.so
echo "PADTO=64 dbg" >> gpurun_out/x2_pad.log
PADTO=64 python tools/time_chain.py 1 2 8 9 >> gpurun_out/x2_pad.log 2>&1
cat gpurun_out/x2_pad.log
