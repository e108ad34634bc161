export PADTO=64
python tools/time_chain.py - ABN_GEMM_ROT=20 ABN_GEMM_ROT=27 ABN_GEMM_ROT=37 ABN_GEMM_ROT=54 ABN_GEMM_GRID=128 ABN_GEMM_GRID=128,ABN_GEMM_ROT=1 ABN_GEMM_GRID=144 > gpurun_out/x3_rot.log 2>&1
cat gpurun_out/x3_rot.log
