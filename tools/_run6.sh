cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputest4.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_gputest4.log
tail -3 gpurun_out/r2b_gputest4.log
timeout 600 python tools/scenario_timing.py > gpurun_out/r2b_scenarios.log 2>&1; cat gpurun_out/r2b_scenarios.log | tail -5
