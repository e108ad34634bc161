python -m pytest tests -m gpu -x -q > gpurun_out/x7_tests.log 2>&1; tail -3 gpurun_out/x7_tests.log
export PADTO=64
for w in 5 7 9 11 14; do WSPLIT=$w python tools/time_chain.py - 2>&1 | sed "s/^-  /WSPLIT=$w/"; done > gpurun_out/x7_merge.log 2>&1
cat gpurun_out/x7_merge.log
python bench.py --no-cpu-baseline > gpurun_out/x7_bench.json 2> gpurun_out/x7_bench.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/x7_bench.json").read().strip().splitlines()[-1])
print(d["train"]["ms_per_step"], d["train"]["tensor_util"], d["train"]["e2e"])
P
ABN_BWD_MERGE=0 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nomerge', d['train']['ms_per_step'])"
