run() { env "$@" python tools/time_step.py "$@"; }
( run ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=18
  run ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=9
  run ABN_BWD_MERGE=0 ABN_WGRAD_SPLIT=14
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=9
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=14
  run ABN_BWD_MERGE=1 ABN_WGRAD_SPLIT=18 ) > gpurun_out/x8_step.log 2>&1
cat gpurun_out/x8_step.log
