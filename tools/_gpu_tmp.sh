cd /root/repo
for v in w18 w22 w24; do
  lib=lqr_obstacles_b200/liblqro${v:+_$v}.so
  echo "== $lib"
  LQRO_LIB=$PWD/$lib python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}"
done
echo "== polling SM sweep (default lib)"
for s in 6 8 10 12; do
  LQRO_HULL_OVERLAP_SMS=$s python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "overlap_sms [0-9]*\|median ms.*total[^}]*}" | tr '\n' ' '; echo
done
for s in 6 8 12 16; do
  LQRO_HULL_OVERLAP_SMS=$s python tools/quick_pairs.py 32x16x16 2>&1 | tail -1 | grep -o "overlap_sms [0-9]*\|median ms.*total[^}]*}" | tr '\n' ' '; echo
done
