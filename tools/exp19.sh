timeout 400 python -m pytest tests/test_gpu_embedder.py -q > gpurun_out/x19_emb.log 2>&1; tail -15 gpurun_out/x19_emb.log
