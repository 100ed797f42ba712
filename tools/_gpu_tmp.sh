cd /root/repo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/final_bench_default.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches'], d['parity_sample_ok'])"
