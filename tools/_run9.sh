cd /root/repo
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2b_bench_n$n.json 2> gpurun_out/r2b_bench_n$n.err; tail -c 300 gpurun_out/r2b_bench_n$n.err
done
python - <<'PY'
import json
for n in (4,8):
    try:
        d=json.loads(open('gpurun_out/r2b_bench_n%d.json'%n).read().strip().split('\n')[-1])
        print(n, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['stages_ms'], d['checksum']['newV_u64'], d['checksum']['status_u64'], d.get('parity_sample_ok'), d['roofline']['frac'], d['roofline'].get('overlap_ms'))
    except Exception as e: print(n,'ERR',e)
PY
