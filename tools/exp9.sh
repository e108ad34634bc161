run() { env "$@" python tools/time_step.py "$@"; }
( run ABN_LIB=$PWD/abnet3_b200/libabnet3_b200_old.so ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=18
  run ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=18
  run ABN_LIB=$PWD/abnet3_b200/libabnet3_b200_old.so ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=18
  run ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=18
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=18
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=24
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=30 ) > gpurun_out/x9_step.log 2>&1
cat gpurun_out/x9_step.log
