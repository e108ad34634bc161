#!/bin/bash
# Experiment build: the same library with the GEMM kernel's debug knobs compiled in
# (-DABN_TC_DEBUG: ABN_GEMM_DBG bit mask, see abn_tc2.cu).  Use: ABN_LIB=$PWD/abnet3_b200/libabnet3_b200_dbg.so
set -e
cd "$(dirname "$0")/../abnet3_b200"
mkdir -p _obj_dbg
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DABN_TC_DEBUG"
for f in abn_capi abn_align abn_nn abn_tc abn_tc2 abn_tc3 abn_fused; do
  nvcc $FL -c csrc/$f.cu -o _obj_dbg/$f.o &
done
wait
nvcc -shared -o libabnet3_b200_dbg.so _obj_dbg/*.o -cudart static
echo built libabnet3_b200_dbg.so
