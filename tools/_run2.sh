cd /root/repo
LQRO_HULL_OVERLAP_SMS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_hull -s 4 -c 1 -f -o gpurun_out/r2b_k_hull python tools/quick_pairs.py 32x16x16 64 > gpurun_out/r2b_ncu_hull.log 2>&1
tail -3 gpurun_out/r2b_ncu_hull.log
