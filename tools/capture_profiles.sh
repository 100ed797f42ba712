#!/bin/bash
# Round-2 evidence for profiles/: run on a B200 box from the repo root (gpurun -- 'bash tools/capture_profiles.sh').
# Every profiled command first runs without ncu (the bench lines below); numbers printed under ncu are never bench values.
# Under ncu kernel launches are serialised: the polling hull launch finds no producer, gives up after 2 ms (k_hull, poll = 1)
# and the wide launch hulls the whole queue -- so k_pairs is captured on the 136 SMs it has in a real step, k_hull at full width.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_gputest_final.log 2>&1; tail -2 $O/r2_gputest_final.log
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --workload swarm4096 --steps 20 --warmup 5 --no-cpu > $O/r2_bench_swarm4096.json 2> $O/r2_bench_swarm4096.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
# the profiling builds (liblqro_prof.so: -DLQRO_PAIR_PROF, liblqro_hullprof.so) travel with the snapshot when they were built in the
# container (python tools/pairprof.py --build); building them here costs a minute of box time each
[ "${SKIP_PROF:-0}" = 1 ] || { [ -f lqr_obstacles_b200/liblqro_prof.so ] || python tools/pairprof.py --build; python tools/pairprof.py > $O/r2_pairprof.log 2>&1; }
[ "${SKIP_HULL:-0}" = 1 ] || { python tools/hullprof.py --build && python tools/hullprof.py > $O/r2_hullprof.log 2>&1; python tools/scenario_timing.py > $O/r2_scenarios.log 2>&1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launch_list_swarm65536.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_launch_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_pairs -s 4 -c 1 -f -o $O/r2_k_pairs python tools/quick_pairs.py 64x32x32 64 > $O/r2_ncu_pairs.log 2>&1
LQRO_HULL_OVERLAP_SMS=0 ncu --set full --import-source on --clock-control none -k regex:k_hull$ -s 4 -c 1 -f -o $O/r2_k_hull python tools/quick_pairs.py 64x32x32 64 > $O/r2_ncu_hull.log 2>&1
tail -c 400 $O/r2_bench_n1.err; ls -la $O | tail -15
