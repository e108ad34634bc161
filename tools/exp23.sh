python -m pytest tests/test_gpu_align.py -q -x > gpurun_out/x23_tests.log 2>&1; tail -3 gpurun_out/x23_tests.log
python tools/time_e2e.py > gpurun_out/x23_e2e.log 2>&1; grep "align_pairs_host\|upload\|copy back" gpurun_out/x23_e2e.log
