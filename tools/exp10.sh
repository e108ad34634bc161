python -m pytest tests -m gpu -x -q > gpurun_out/x10_tests.log 2>&1; tail -2 gpurun_out/x10_tests.log
PARTS=1 python tools/time_step.py > gpurun_out/x10_parts.log 2>&1; cat gpurun_out/x10_parts.log
python bench.py > gpurun_out/x10_bench.json 2> gpurun_out/x10_bench.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/x10_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["train"]["ms_per_step"], d["train"]["tensor_util"], d["train"]["e2e"]["value"], d["cpu_baseline"])
P
