#!/bin/bash
# Round-2 (second half) checks on one B200: tests, headline bench, launch list and one ncu --set full
# capture of the training kernels (reduced on the box: the .ncu-rep is too large to travel).
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; tail -3 gpurun_out/r2b_tests.log
python bench.py > gpurun_out/r2b_c3_n1.json 2> gpurun_out/r2b_c3_n1.err; tail -c 300 gpurun_out/r2b_c3_n1.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2b_c3_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["train_step"], d["roofline"]["frac"], d["roofline"]["kernels_us_alone"], d["e2e"]["value"], d["clocks"])
P
python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2b_c4_n1.json 2> gpurun_out/r2b_c4_n1.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2b_c4_n1.json").read().strip().splitlines()[-1])
print("C4", d["value"], d["ms_per_step"], d["train_step"], d["roofline"]["kernels_us_alone"])
P
python bench.py --steps 1 --warmup 3 --epoch-batches 20 --no-e2e --no-cpu-baseline --no-align-leg --no-kernel-times > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --epoch-batches 20 --no-e2e --no-cpu-baseline --no-align-leg --no-kernel-times > gpurun_out/r2b_ncu_list.log 2>&1
python tools/prof_r2.py > gpurun_out/r2b_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:mlp_chain|tc_group|gather_bf16|optimizer_fused" -c 15 \
    -o /tmp/r2b_train python tools/prof_r2.py > gpurun_out/r2b_ncu_full.log 2>&1
python tools/ncu_extract.py /tmp/r2b_train.ncu-rep gpurun_out/r2b_train_raw.csv
