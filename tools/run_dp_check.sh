python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/x16_dp$1.log 2>&1
grep -v "^\[W\|NCCL version\|^$" gpurun_out/x16_dp$1.log | tail -14
