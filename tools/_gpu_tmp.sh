cd /root/repo
for rep in 1 2; do
for v in nolead ""; do
  lib=lqr_obstacles_b200/liblqro${v:+_$v}.so
  echo -n "${v:-lead} 8192: "
  LQRO_LIB=$PWD/$lib python tools/quick_pairs.py 32x16x16 2>&1 | tail -1 | grep -o "'polling_hull': [0-9.]*\|median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
done; done
for v in nolead ""; do
  lib=lqr_obstacles_b200/liblqro${v:+_$v}.so
  echo -n "${v:-lead} 4096x32: "
  LQRO_LIB=$PWD/$lib python tools/quick_pairs.py 16x16x16 32 2>&1 | tail -1 | grep -o "'polling_hull': [0-9.]*\|median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
done
echo -n "lead 65536: "; python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum [-0-9.]*" | tr '\n' ' '; echo
