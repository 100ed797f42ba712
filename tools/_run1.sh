set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputest1.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_gputest1.log
( for dims in 64x32x32 32x16x16 ; do for sms in 0 8 12 16; do for fs in 1 2; do LQRO_HULL_OVERLAP_SMS=$sms LQRO_FIRST_SLOTS=$fs timeout 120 python tools/quick_pairs.py $dims 64; done; done; done ) > gpurun_out/r2b_poll_sweep1.log 2>&1
tail -3 gpurun_out/r2b_gputest1.log; cat gpurun_out/r2b_poll_sweep1.log | cut -c1-400
