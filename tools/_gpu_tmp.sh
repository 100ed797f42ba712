cd /root/repo
echo "== 16384 agents (one rank of four): polling SMs"
for s in 12 16 20 24 28; do
  echo -n "sms $s  "
  LQRO_HULL_OVERLAP_SMS=$s python tools/quick_pairs.py 32x32x16 2>&1 | tail -1 | grep -o "'polling_hull': [0-9.]*\|median ms.*total[^}]*}\|'inside': [0-9]*" | tr '\n' ' '; echo
done
echo "== 32768 agents (one rank of two)"
for s in 12 16 20; do
  echo -n "sms $s  "
  LQRO_HULL_OVERLAP_SMS=$s python tools/quick_pairs.py 32x32x32 2>&1 | tail -1 | grep -o "'polling_hull': [0-9.]*\|median ms.*total[^}]*}\|'inside': [0-9]*" | tr '\n' ' '; echo
done
