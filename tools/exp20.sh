python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/x20_smoke.log 2>&1; tail -3 gpurun_out/x20_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/x20_tests.log 2>&1; tail -3 gpurun_out/x20_tests.log
