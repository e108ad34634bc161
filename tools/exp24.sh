python bench.py > gpurun_out/r1r_bench.json 2> gpurun_out/r1r_bench.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r1r_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["e2e"], d["train"]["ms_per_step"], d["train"]["tensor_util"], d["train"]["e2e"]["value"], d["clocks"], d["cpu_baseline"]["value"])
P
python -m pytest tests -m gpu -x -q > gpurun_out/r1r_tests.log 2>&1; tail -2 gpurun_out/r1r_tests.log
