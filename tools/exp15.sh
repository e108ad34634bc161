timeout 300 python -m pytest tests/test_gpu_tc3.py -q > gpurun_out/x15_tests.log 2>&1; tail -3 gpurun_out/x15_tests.log
timeout 200 python tools/time_fused.py > gpurun_out/x15_time.log 2>&1; tail -4 gpurun_out/x15_time.log
run() { env "$@" python tools/time_step.py "$@"; }
( run ABN_FWD_FUSED=0 ABN_BWD_FUSED=0
  run ABN_FWD_FUSED=1 ABN_BWD_FUSED=0
  run ABN_FWD_FUSED=0 ABN_BWD_FUSED=1
  run ABN_FWD_FUSED=1 ABN_BWD_FUSED=1
  run ABN_FWD_FUSED=1 ABN_BWD_FUSED=1 ABN_NO_PDL=1
  run ABN_FWD_FUSED=1 ABN_BWD_FUSED=1 ABN_WGRAD_SPLIT=7
  run ABN_FWD_FUSED=1 ABN_BWD_FUSED=1 ABN_WGRAD_SPLIT=12 ) > gpurun_out/x15_step.log 2>&1
cat gpurun_out/x15_step.log
