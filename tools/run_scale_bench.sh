N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r1q_bench$N.json 2> gpurun_out/r1q_bench$N.err
python - <<P
import json
d=json.loads(open("gpurun_out/r1q_bench$N.json").read().strip().splitlines()[-1])
print("N=$N", d["value"], (d.get("e2e") or {}).get("value"), d["train"]["value"], d["train"]["ms_per_step"], d["train"]["e2e"]["value"])
P
