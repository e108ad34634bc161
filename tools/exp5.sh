python -m pytest tests -m gpu -x -q > gpurun_out/x5_tests.log 2>&1; tail -3 gpurun_out/x5_tests.log
export PADTO=64
for w in 6 9 12 18 24 36; do WSPLIT=$w python tools/time_chain.py - 2>&1 | sed "s/^-/WSPLIT=$w/"; done > gpurun_out/x5_wsplit.log 2>&1
cat gpurun_out/x5_wsplit.log
python bench.py --no-cpu-baseline > gpurun_out/x5_bench.json 2> gpurun_out/x5_bench.err; tail -c 1500 gpurun_out/x5_bench.json
