cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k riccati 2>&1 | tail -2
timeout 300 python tools/quick_riccati.py > gpurun_out/r2b_riccati2.json 2>gpurun_out/r2b_riccati2.err; cat gpurun_out/r2b_riccati2.json; tail -3 gpurun_out/r2b_riccati2.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_riccati -s 1 -c 1 -f -o gpurun_out/r2b_k_riccati2 python tools/quick_riccati.py > gpurun_out/r2b_ncu_riccati2.log 2>&1; tail -2 gpurun_out/r2b_ncu_riccati2.log
