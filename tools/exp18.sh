python tools/prof_step.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:mlp_chain|tc_group" -s 9 -c 3 -f -o gpurun_out/r1q_fused python tools/prof_step.py > gpurun_out/ncu_q2.log 2>&1
tail -5 gpurun_out/ncu_q2.log; ls -la gpurun_out/r1q_fused.ncu-rep
