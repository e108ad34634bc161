python -m pytest tests -m gpu -x -q > gpurun_out/r1q_tests.log 2>&1; tail -3 gpurun_out/r1q_tests.log
python bench.py > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r1q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["train"]["ms_per_step"], d["train"]["tensor_util"], d["train"]["e2e"]["value"], d["clocks"])
P
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1q_ref.json 2>/dev/null; tail -c 400 gpurun_out/r1q_ref.json
python tools/prof_step.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1q_launches_step.csv python tools/prof_step.py > gpurun_out/ncu_q1.log 2>&1
tail -14 gpurun_out/r1q_launches_step.csv | cut -c1-200
