cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_edges.py tests/test_gpu_trajectory.py -m gpu -x -q > gpurun_out/r2b_gputest6.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_gputest6.log
tail -3 gpurun_out/r2b_gputest6.log
( for dims in 64x32x32 32x16x16 ; do for sms in 0 12; do LQRO_HULL_OVERLAP_SMS=$sms timeout 120 python tools/quick_pairs.py $dims 64; done; done ) > gpurun_out/r2b_poll_sweep4.log 2>&1
cut -c1-330 gpurun_out/r2b_poll_sweep4.log
