cd /root/repo
timeout 150 python tools/quick_e2e.py > gpurun_out/r4_e2e.log 2>&1; echo "e2e rc $?"
grep -o "^[a-z_]* [a-z ]*e2e ms/step median [0-9.]*\|identical.*" gpurun_out/r4_e2e.log
timeout 100 python tools/quick_e2e.py 32x32x32 > gpurun_out/r4_e2e_32768.log 2>&1; echo "e2e rc $?"
grep -o "^[a-z_]* [a-z ]*e2e ms/step median [0-9.]*\|'hull': [0-9.]*\|identical.*" gpurun_out/r4_e2e_32768.log
timeout 300 python -m pytest tests/test_gpu_headline.py tests/test_gpu_edges.py -m gpu -x -q -k "launch_shape or chunked" 2>&1 | tail -8 | tee gpurun_out/r4_tests.log
