export PADTO=64
for w in 4 5 7 8 9 10 11 14; do WSPLIT=$w python tools/time_chain.py - 2>&1 | sed "s/^-/WSPLIT=$w/" | cut -c1-10,150-; done > gpurun_out/x6_wsplit.log 2>&1
cat gpurun_out/x6_wsplit.log
