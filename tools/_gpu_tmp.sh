cd /root/repo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4_bench_n2.json 2> gpurun_out/r4_bench_n2.err; echo "rc $?"
tail -3 gpurun_out/r4_bench_n2.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4_bench_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['neighbour_table_upload'][:30], d['checksum']['newV_u64'], d['checksum']['status_u64'], d['parity_sample_ok'], d['config']['parallelism'][:80])
P
