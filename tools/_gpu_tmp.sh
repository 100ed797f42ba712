cd /root/repo; O=gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/r4_smoke.log
timeout 60 python bench.py --workload swarm4096 --steps 20 --warmup 5 --no-cpu > $O/r4_bench_swarm4096.json 2> $O/r4_bench_swarm4096.err; echo "rc $?"
