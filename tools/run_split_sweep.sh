run() { env "$@" python tools/time_step.py "$@"; }
( run ABN_WGRAD_SPLIT=9
  run ABN_WGRAD_SPLIT=5
  run ABN_WGRAD_SPLIT=4
  run ABN_WGRAD_SPLIT=6
  run ABN_WGRAD_SPLIT=10 ) 2>&1 | grep step
