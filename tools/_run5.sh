cd /root/repo
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2b_multirank.log 2>&1; echo "rc $?" >> gpurun_out/r2b_multirank.log; tail -3 gpurun_out/r2b_multirank.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 600 gpurun_out/r2b_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -c 600 gpurun_out/r2b_bench_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_n%d.json'%n).read().strip().split('\n')[-1])
        print(n, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['stages_ms'], d['checksum']['newV_u64'], d['checksum']['status_u64'], d.get('parity_sample_ok'), d['roofline']['frac'])
    except Exception as e: print(n,'ERR',e)
PY
