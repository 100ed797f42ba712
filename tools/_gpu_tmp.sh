cd /root/repo; O=gpurun_out
timeout 45 python tools/quick_e2e.py 32x32x16 64 30 > $O/r4_e2e_16384.log 2>&1; echo "rc $?"
grep -o "^[a-z_]* [a-z ]*e2e ms/step median [0-9.]* min [0-9.]*\|'hull': [0-9.]*\|identical.*" $O/r4_e2e_16384.log
