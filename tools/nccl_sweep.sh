for cfg in "default" "NCCL_ALGO=NVLS" "NCCL_ALGO=Tree" "NCCL_PROTO=LL" "NCCL_PROTO=LL128" "NCCL_ALGO=Ring NCCL_PROTO=LL128" "NCCL_NVLS_ENABLE=1 NCCL_ALGO=NVLSTree"; do
  echo "== $cfg"
  if [ "$cfg" = "default" ]; then E=""; else E="$cfg"; fi
  env $E ABN_DP_P2P=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 50000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['train']['ms_per_step'])"
done
