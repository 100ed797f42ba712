cd /root/repo
timeout 300 python tools/quick_riccati.py > gpurun_out/r2b_riccati.json 2>gpurun_out/r2b_riccati.err; cat gpurun_out/r2b_riccati.json; tail -3 gpurun_out/r2b_riccati.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_riccati -s 1 -c 1 -f -o gpurun_out/r2b_k_riccati python tools/quick_riccati.py > gpurun_out/r2b_ncu_riccati.log 2>&1; tail -2 gpurun_out/r2b_ncu_riccati.log
