#!/usr/bin/env python
"""bench.py -- agent-neighbour LQR-obstacle constructions per second (BASELINE.json metric).

A "step" is one pass of the hot path (obstacle creation -> half-plane creation -> new-velocity LP, with the
state all-gather when N > 1) over the synthetic 65,536-agent swarm, K = 64 neighbours, H = 20 horizon
samples, P = 50 ellipsoid samples (BASELINE.json configs[4]); agents are partitioned over the N GPUs
(strong scaling: the swarm is fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload swarm65536|swarm4096]
  N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

ours      : value = pairs / device time with everything resident in HBM (CUDA events on the launching
            stream, L2 flushed between timed steps, max over ranks); e2e = the same through the C-ABI call
            lqro_step_sharded() with pinned HOST buffers (H2D of state/vpref/neighbours and D2H of the new
            velocities inside the timed region).
            N = 1 also reports, outside `value` / `e2e`: `sim_update` -- the second half of the reference's time step
            (LQR feedback, propagate, Kalman filter: lqro_sim_update_device) for the same agents, per-kernel ms, the
            reference's own update loop on one host core beside it, and `sim_update.timestep`: whole time steps on
            resident state (lqro_timestep_device) in closed loop.
reference : the reference's own CPU implementation (oracle/_ref, built from /root/reference; else the
            oracle port) on the host cores, sharded by process, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "swarm65536": dict(dims=(64, 32, 32), K=64, H=20, P=50),
    "swarm4096": dict(dims=(16, 16, 16), K=32, H=20, P=50),
}
METRIC = "agent_neighbour_lqr_obstacle_constructions_per_sec"
UNIT = "pairs/s"


def f_alg(H, P):
    """deterministic FP64 flop per pair: H*(96 + 32*P)"""
    return H * (96 + 32 * P)


def b_alg(K):
    """compulsory HBM bytes per pair: 160 + 176/K"""
    return 160.0 + 176.0 / K


def load_dynamics():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    return g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"]


def make_workload(name):
    from lqr_obstacles_b200 import swarm

    w = WORKLOADS[name]
    x, vpref = swarm.lattice_swarm(w["dims"])
    nbr = swarm.knn(x[:, :3], w["K"])
    return x, vpref, nbr, w


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
_REF = None
_WL = None  # (A, B, L, E, NF, T, sph, x, vpref, nbr, H): set before the pool forks, inherited by the workers


def _ref_init():
    global _REF
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _libs import REF_SO, Oracle, Reference

    if os.path.exists(REF_SO):
        _REF = ("reference", Reference())
    else:
        _REF = ("port", Oracle())


def _ref_chunk(agents):
    """one shard of agents through the reference's own per-pair code + calculateNewV, looped in C++
    (ref_step_agents in oracle/ref_harness.cpp; lqro_o_step for the port)."""
    kind, R = _REF
    A, B, L, E, NF, T, sph, x, vpref, nbr, H = _WL
    rows = np.ascontiguousarray(nbr[agents])
    if kind == "reference":
        import ctypes
        R.lib.ref_stage_seconds(None, None, 1)
        npairs, newV, _ = R.step_agents(A, B, L, E, x, vpref, rows, agents, H, 30.0)
        st = np.zeros(4)
        nh = ctypes.c_long()
        R.lib.ref_stage_seconds(st.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nh), 1)
        return int(npairs), newV, st, int(nh.value)
    xs = x  # the port takes the same global state; restrict the neighbour table to the shard
    full = np.full((len(x), rows.shape[1]), -1, dtype=np.int32)
    full[agents] = rows
    newV = R.step(NF, T, sph, xs, vpref, full, nthreads=1)[0][agents]
    return int((rows >= 0).sum()), newV, np.zeros(4), 0


class CpuBaseline:
    """The reference CPU path on the host cores.  Sharded by PROCESS (the reference's GJK scratch and world
    constants are file-static/global); every process loads its own copy of the library."""

    def __init__(self, workload, nproc=None):
        import multiprocessing as mp

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _libs import REF_SO, Oracle, build_oracle

        build_oracle()
        self.kind = "reference" if os.path.exists(REF_SO) else "port"
        self.x, self.vpref, self.nbr, self.w = workload
        self.A, self.B, self.L, self.E = load_dynamics()
        O = Oracle()
        self.NF, self.T, _ = O.horizon(self.A, self.B, self.L, self.E, self.w["H"])
        self.sph = O.spheres(self.w["P"])
        self.cores = nproc or len(os.sched_getaffinity(0))
        global _WL
        _WL = (self.A, self.B, self.L, self.E, self.NF, self.T, self.sph, self.x, self.vpref, self.nbr, self.w["H"])
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_ref_init)

    def run(self, n_agents_sample, seed=0):
        N = len(self.x)
        rng = np.random.default_rng(seed)
        agents = np.sort(rng.choice(N, size=min(n_agents_sample, N), replace=False))
        tasks = [s.astype(np.int32) for s in np.array_split(agents, self.cores * 4) if len(s)]
        t0 = time.perf_counter()
        res = self.pool.map(_ref_chunk, tasks)
        dt = time.perf_counter() - t0
        pairs = sum(r[0] for r in res)
        # CPU seconds (summed over processes) in the reference's own categories, SURVEY 8(d)
        st = np.sum([r[2] for r in res], axis=0)
        self.stages = {"unit": "cpu-us per pair (summed over processes / pairs)",
                       "obstacle_creation": 1e6 * st[0] / max(pairs, 1), "half_plane_creation": 1e6 * st[1] / max(pairs, 1),
                       "new_velocity": 1e6 * st[2] / max(pairs, 1), "hull_branch_share_of_half_plane": st[3] / max(st[1], 1e-30),
                       "hulls": int(sum(r[3] for r in res)),
                       "hull_ms_each": 1e3 * st[3] / max(1, sum(r[3] for r in res)),
                       "note": "hull = in-process qhull-order quickhull + the reference's text round trip, no process spawn"}
        return pairs, dt, len(agents)

    def close(self):
        self.pool.terminate()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload)
    cpu = CpuBaseline(wl)
    sample = args.ref_agents
    for _ in range(args.warmup):
        cpu.run(max(cpu.cores, sample // 8))
    pairs = 0
    dt = 0.0
    for s in range(args.steps):
        p, d, na = cpu.run(sample, seed=s)
        pairs += p
        dt += d
    cpu.close()
    value = pairs / dt
    N = len(wl[0])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d agents, K=%d, H=%d, P=%d" % (args.workload, N, wl[3]["K"], wl[3]["H"], wl[3]["P"]),
                   "sample": "%d randomly chosen agents per step (all their K pairs + LP)" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "stages": getattr(cpu, "stages", None),
                         "sample": "%d agents x K pairs per step, %d steps, process-sharded" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def sim_update_arm(ctx, x, L, E, stream, args):
    """The simulator-form update (k_sim_flow, k_sim_noise, k_sim_filter) over every agent of the swarm: estimates = the swarm's states, identity attitudes, P = 1e-9 I, synthetic
    standard normals (numpy; the kernel's cost does not depend on their values), gains from the golden vectors.  Timed
    with CUDA events around each launch (lqro_sim_last_ms); beside it the reference's own update loop on one host core
    (oracle/_ref, 32 quads) when it is available."""
    import torch
    from lqr_obstacles_b200.api import sim_default_model
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    n = len(x)
    ctx.sim_set_model(sim_default_model(L=L, E=E, Lh=g["dyn_Lh"], Eh=g["dyn_Eh"]))
    rng = np.random.default_rng(1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    eye3, p0 = np.eye(3).ravel(), (1e-9 * np.eye(16)).ravel()
    dx, dxt = dev(x), dev(x)
    dR, dRT = dev(np.tile(eye3, (n, 1))), dev(np.tile(eye3, (n, 1)))
    dP = dev(np.tile(p0, (n, 1)))
    dnv = dev(x[:, 3:6] + 0.1 * rng.normal(size=(n, 3)))
    dpg = dev(-x[:, 0:3])
    dnr = dev(rng.normal(size=(n, 22)))
    dvg, du = torch.zeros(n, 3, dtype=torch.float64, device="cuda"), torch.zeros(n, 4, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ms = []
    for it in range(2 + max(1, min(args.steps, 5))):
        ctx.sim_update_device(n, dx.data_ptr(), dR.data_ptr(), dP.data_ptr(), dxt.data_ptr(), dRT.data_ptr(), dnv.data_ptr(),
                              dpg.data_ptr(), dnr.data_ptr(), dvg.data_ptr(), du.data_ptr())
        ms.append(ctx.sim_last_ms(parts=True))
    ms = np.array(ms[2:])
    k_ms = float(ms[:, 0].mean())
    arm = {"kernels": {"k_sim_flow": float(ms[:, 1].mean()), "k_sim_noise": float(ms[:, 2].mean()), "k_sim_filter": float(ms[:, 3].mean())},
           "agents": n, "ms_per_update": k_ms, "agents_per_s": n / (k_ms * 1e-3),
           "finite": bool(torch.isfinite(dx).all().item() and torch.isfinite(dP).all().item()),
           "api": "lqro_sim_update_device (state resident; standard normals supplied by the host, see lqro_msvc_normals)"}
    # whole time step resident on the device (lqro_timestep_device: avoidance step on the context's state + the update;
    # the neighbour table is the resident one): a closed loop of a few steps, CUDA events on the launching stream
    ts = []
    for it in range(2 + max(1, min(args.steps, 5))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.timestep_device(dR.data_ptr(), dP.data_ptr(), dxt.data_ptr(), dRT.data_ptr(), dpg.data_ptr(), dnr.data_ptr(), dvg.data_ptr())
        e1.record(stream)
        stream.synchronize()
        ts.append(e0.elapsed_time(e1))
    arm["timestep"] = {"ms_per_step": float(np.mean(ts[2:])), "api": "lqro_timestep_device (MAIN:1373-1417 for every agent: k_pairs, k_hull, "
                       "k_new_velocity, k_sim_control, k_sim_flow, k_sim_noise, k_sim_filter; state resident)",
                       "gpu_launches_per_step": 7, "finite": bool(torch.isfinite(dvg).all().item())}
    if not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from _libs import Reference, RefSim
            ref = Reference()
            m = 32
            ang = np.arange(m) * 2 * np.pi / m
            starts = np.stack([12 * np.cos(ang), 12 * np.sin(ang), np.zeros(m)], 1)
            sim = RefSim(ref, starts, -starts, seed=1)
            sim.set_newv(np.zeros((m, 3)))
            sim.update()
            t0 = time.perf_counter()
            for _ in range(20):
                sim.update()
            d = time.perf_counter() - t0
            sim.close()
            arm["cpu_reference"] = {"agents_per_s": 20 * m / d, "cores": 1, "kind": "reference",
                                    "sample": "the reference's own update loop (oracle/_ref ref_sim_update), 32 quads x 20 steps, %.2f s" % d}
        except Exception as e:  # oracle/_ref not present on this box
            arm["cpu_reference"] = {"unavailable": str(e)[:120]}
    return arm


def run_ours(args):
    import torch
    import torch.distributed as dist

    from lqr_obstacles_b200 import Lqro

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    x, vpref, nbr, w = make_workload(args.workload)
    N, K, H, P = len(x), w["K"], w["H"], w["P"]
    assert N % world == 0
    NL = N // world
    lo = rank * NL
    A, B, L, E = load_dynamics()
    ctx = Lqro(device=local_rank, n_agents=N, local_begin=lo, n_local=NL, max_neighbors=K, horizon=H, n_samples=P)
    ctx.set_dynamics(A, B, L, E)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(Lqro.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- device-resident arm: this rank's block of the state is already in the context's buffer
    x_loc = np.ascontiguousarray(x[lo:lo + NL])
    vp_loc = np.ascontiguousarray(vpref[lo:lo + NL])
    nbr_loc = np.ascontiguousarray(nbr[lo:lo + NL])
    d_vpref = torch.from_numpy(vp_loc).cuda()
    d_nbr = torch.from_numpy(nbr_loc).cuda()
    d_newV = torch.zeros(NL, 3, dtype=torch.float64, device="cuda")
    ctx.upload_local_state(x_loc)
    ctx.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def one_step():
        ctx.allgather_state()  # no-op on one GPU
        ctx.step_device(None, d_vpref.data_ptr(), d_nbr.data_ptr(), d_newV.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = np.zeros(4)
    barrier()
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            if world > 1:
                dist.barrier()  # line the ranks up before every timed step: host-side skew is not step time
            flush.zero_()  # L2 flush, outside the per-step events
            ev[s][0].record(stream)
            one_step()
            ev[s][1].record(stream)
            stream.synchronize()
            t = ctx.last_step_ms()
            kern_ms += np.array([t["pairs"], t["hull"], t["lp"], t["total"]])
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    kern_ms /= args.steps
    counters = ctx.counters()

    # ---- end-to-end arm: host buffers (pinned), copies inside the timed region
    hx = torch.from_numpy(x_loc).pin_memory()
    hv = torch.from_numpy(vp_loc).pin_memory()
    hn = torch.from_numpy(nbr_loc).pin_memory()
    hout = torch.zeros(NL, 3, dtype=torch.float64).pin_memory()
    for _ in range(max(1, args.warmup // 2)):
        ctx.step_sharded(hx.data_ptr(), hv.data_ptr(), hn.data_ptr(), hout.data_ptr())
    barrier()
    t0 = time.perf_counter()
    e2e_ev0, e2e_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ev0.record(stream)
    for s in range(args.steps):
        ctx.step_sharded(hx.data_ptr(), hv.data_ptr(), hn.data_ptr(), hout.data_ptr())
    e2e_ev1.record(stream)
    stream.synchronize()
    e2e_ms = e2e_ev0.elapsed_time(e2e_ev1)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    checksum = float(hout.sum())
    e2e_launches = ctx.counters()["launches"]  # pair kernel per uploaded neighbour block + hull + LP

    # ---- the same call with the neighbour table rebuilt on the device every step (LQRO_FLAG_DEVICE_KNN, nbr = NULL):
    # informative; the headline e2e keeps the host-provided lists, like the reference arm
    knn_arm = None
    if world == 1:
        from lqr_obstacles_b200.api import FLAG_DEVICE_KNN
        ctx2 = Lqro(device=local_rank, n_agents=N, max_neighbors=K, horizon=H, n_samples=P, flags=FLAG_DEVICE_KNN)
        ctx2.set_dynamics(A, B, L, E)
        ctx2.set_stream(stream.cuda_stream)
        hout2 = torch.zeros(NL, 3, dtype=torch.float64).pin_memory()
        for _ in range(2):
            ctx2.step_sharded(hx.data_ptr(), hv.data_ptr(), None, hout2.data_ptr())
        stream.synchronize()
        t0 = time.perf_counter()
        for s in range(args.steps):
            ctx2.step_sharded(hx.data_ptr(), hv.data_ptr(), None, hout2.data_ptr())
        stream.synchronize()
        k_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        knn_arm = {"value": N * K / (k_ms * 1e-3), "unit": UNIT, "ms_per_step": k_ms,
                   "h2d_bytes_per_step": int(x_loc.nbytes + vp_loc.nbytes), "gpu_launches_per_step": 9,
                   "neighbour_table_identical": bool(np.array_equal(ctx2.neighbors(), nbr_loc)),
                   "checksum": float(hout2.sum()), "api": "lqro_step_sharded(nbr=NULL) with LQRO_FLAG_DEVICE_KNN"}
        ctx2.close()

    # ---- the second half of the reference's time step (MAIN:1408-1417: LQR feedback, propagate, Kalman filter) for the same
    # swarm on the device (lqro_sim_update_device): informative, not part of `value`
    sim_arm = None
    if world == 1:
        sim_arm = sim_update_arm(ctx, x, L, E, stream, args)

    t_dev = torch.tensor([dev_ms, e2e_ms, e2e_wall_ms, kern_ms[0]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_wall_ms, pair_ms = [float(v) for v in t_dev.cpu()]
    total_pairs = N * K
    value = total_pairs * args.steps / (dev_ms * 1e-3)
    e2e_value = total_pairs * args.steps / (max(e2e_ms, e2e_wall_ms) * 1e-3)

    if rank == 0:
        fp64_peak, fp32_peak = ctx.measure_peaks()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        pairs_per_launch = NL * K
        pair_s = pair_ms * 1e-3
        achieved_tf = pairs_per_launch * f_alg(H, P) / pair_s / 1e12
        # DRAM traffic and executed-FP64 counts of k_pairs come from the committed ncu capture of this workload
        # (profiles/rNN_k_pairs_<workload>_ncu.json, written by tools/ from the .ncu-rep); never measured under ncu here
        ncu = None
        import glob
        caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_pairs_%s_ncu.json" % args.workload)))
        if caps and world == 1:
            try:
                ncu = json.load(open(caps[-1]))
                ncu["file"] = os.path.relpath(caps[-1], ROOT)
            except Exception:
                ncu = None
        roofline = {
            "bound": "fp64", "kernel": "k_pairs", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak,
            "traffic": (ncu["traffic_bytes"] if ncu and ncu.get("pairs_per_launch") == pairs_per_launch else None),
            "algorithmic_bytes_per_launch": pairs_per_launch * b_alg(K),
            "executed": ({"source": ncu["file"], "fp64_flop_per_launch": ncu["executed_fp64_flop"],
                          "fp64_pipe_pct": ncu["fp64_pipe_pct"], "issue_active_pct": ncu["issue_active_pct"],
                          "note": "exact pruning skips horizon steps, so executed flop < algorithmic flop (F_alg counts every step)"}
                         if ncu else None),
            "algorithmic_flop_per_pair": f_alg(H, P), "pairs_per_launch": pairs_per_launch, "kernel_ms": pair_ms,
            "peak_source": "FP64 CUDA-core pipe measured in this run (DFMA chains, lqro_measure_fp64_peak); "
                           "MEASURED_PEAKS.json has no FP64 figure",
            "hbm": {"achieved": pairs_per_launch * b_alg(K) / pair_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": pairs_per_launch * b_alg(K) / pair_s / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
            "fp32_peak_tflops": fp32_peak,
            "step_kernels_ms": {"k_pairs": kern_ms[0], "k_hull": kern_ms[1], "k_new_velocity": kern_ms[2], "sum": kern_ms[3]},
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu:
            cpu = CpuBaseline((x, vpref, nbr, w))
            cpu.run(cpu.cores)  # warm the pool
            p, d, na = cpu.run(args.cpu_agents)
            cpu.close()
            cpu_baseline = {"value": p / d, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "stages": getattr(cpu, "stages", None),
                            "sample": "%d randomly chosen agents (x K=%d pairs + LP) of the same swarm, process-sharded, %.1f s" % (na, K, d)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d agents, K=%d, H=%d, P=%d, jittered 4 m lattice, seed 20261018" % (args.workload, N, K, H, P),
                       "parallelism": "agents sharded x%d, ncclAllGather of the 128-B state per step" % world,
                       "l2": "256 MiB buffer written between timed steps (outside the per-step events)",
                       "pairs_per_step": total_pairs},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * (x_loc.nbytes + vp_loc.nbytes + nbr_loc.nbytes)),
                    "d2h_bytes_per_step": int(world * NL * 24), "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps,
                    "api": "lqro_step_sharded (host pinned buffers; neighbour table uploaded in blocks overlapped with the pair kernel)",
                    "gpu_launches_per_step": e2e_launches, "checksum": checksum, "device_knn": knn_arm},
            "gpu_launches": 3 * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "counters": counters,
            "sim_update": sim_arm,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="swarm65536", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-agents", type=int, default=32768, help="agents in the cpu_baseline sample (~10-30 s of CPU work)")
    ap.add_argument("--ref-agents", type=int, default=8192, help="agents per step of --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
