python -m pytest tests/test_gpu_tc.py -q > gpurun_out/x21_tc.log 2>&1; tail -15 gpurun_out/x21_tc.log
