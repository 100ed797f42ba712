cd /root/repo; O=gpurun_out
for v in prev cur prev cur; do
  lib=lqr_obstacles_b200/liblqro.so; [ $v = prev ] && lib=lqr_obstacles_b200/liblqro_prev.so
  echo -n "$v 65536: "; LQRO_LIB=$PWD/$lib timeout 60 python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
done | tee $O/r4_ab2.log
timeout 260 python -m pytest tests -m gpu -x -q > $O/r4_gputest.log 2>&1; tail -2 $O/r4_gputest.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r4_bench_n1.json 2> $O/r4_bench_n1.err; echo "bench rc $?"
timeout 120 ncu --set full --import-source on --clock-control none -k regex:k_pairs -s 4 -c 1 -f -o $O/r4_k_pairs python tools/quick_pairs.py 64x32x32 64 > $O/r4_ncu_pairs.log 2>&1; echo "ncu rc $?"
