cd /root/repo
python tools/quick_pairs.py 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum.*"  | tr '\n' ' '; echo
python tools/quick_pairs.py 32x16x16 2>&1 | tail -1 | grep -o "median ms.*total[^}]*}\|checksum.*"  | tr '\n' ' '; echo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3
