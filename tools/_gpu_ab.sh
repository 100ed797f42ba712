cd /root/repo; O=gpurun_out
for rep in 1 2; do for v in prev cur; do
  lib=lqr_obstacles_b200/liblqro.so; [ $v = prev ] && lib=lqr_obstacles_b200/liblqro_prev.so
  echo -n "$v 65536: "; LQRO_LIB=$PWD/$lib timeout 60 python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
done; done | tee $O/r4_ab.log
for v in prev cur; do
  lib=lqr_obstacles_b200/liblqro.so; [ $v = prev ] && lib=lqr_obstacles_b200/liblqro_prev.so
  echo -n "$v 8192: "; LQRO_LIB=$PWD/$lib timeout 60 python tools/quick_pairs.py 32x16x16 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
done | tee -a $O/r4_ab.log
