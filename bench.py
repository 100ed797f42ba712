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


def reference_dynamics():
    """A, B, L, E as the reference's own controlMatrices computed them (golden vectors generated from oracle/_ref): the
    dynamics of the --impl reference arm, and the yardstick the device-computed gains are compared with."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    return g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"]


def device_dynamics(ctx):
    """controlMatrices (MAIN:490-544) on the device: lqro_linearize_discretize at the hover point + lqro_solve_gains.
    Returns A, B, L, E, Lh, Eh and the largest relative deviation of L, E from the reference's."""
    from lqr_obstacles_b200.api import sim_default_model

    m = sim_default_model()
    ctx.sim_set_model(m)
    u0 = float(m.u_goal[0])
    xs = np.zeros(16)
    xs[12:] = u0
    A, B, c = ctx.linearize_discretize(xs, np.eye(3), np.full(4, u0))
    L, E, Lh, Eh = ctx.solve_gains(A, B)
    rA, rB, rL, rE = reference_dynamics()
    dev = max(float(np.abs(L[0] - rL).max() / np.abs(rL).max()), float(np.abs(E[0] - rE).max() / np.abs(rE).max()),
              float(np.abs(A[0] - rA).max() / np.abs(rA).max()), float(np.abs(B[0] - rB).max() / np.abs(rB).max()))
    return A[0], B[0], L[0], E[0], Lh[0], Eh[0], dev


def make_workload(name):
    from lqr_obstacles_b200 import swarm

    w = WORKLOADS[name]
    x, vpref = swarm.lattice_swarm(w["dims"])
    nbr = swarm.knn(x[:, :3], w["K"])
    return x, vpref, nbr, w


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  nvidia-smi needs a few
    hundred ms to come up and a timed region of a few steps lasts tens of ms, so it is started before the warm-up and polls every
    20 ms; begin()/end() bracket the timed region and only the samples inside it (plus the one right before it, taken
    under the warm-up load) are used."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc, self.t0, self.t1 = gpu_index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.03]   # a sample printed <= 30 ms after the region was taken inside it
        before = [r for ts, r in self.rows if ts < t0]
        rows = inside + before[-1:]                                     # plus the last one of the warm-up (same load)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside),
                "timed_region_ms": (None if self.t1 is None or self.t0 is None else 1e3 * (self.t1 - self.t0))}


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
        cb_s = R.lib.ref_hull_callback_seconds(1) if hasattr(R.lib, "ref_hull_callback_seconds") else float("nan")
        return int(npairs), newV, st, int(nh.value), cb_s
    xs = x  # the port takes the same global state; restrict the neighbour table to the shard
    full = np.full((len(x), rows.shape[1]), -1, dtype=np.int32)
    full[agents] = rows
    newV = R.step(NF, T, sph, xs, vpref, full, nthreads=1)[0][agents]
    return int((rows >= 0).sum()), newV, np.zeros(4), 0, float("nan")


class CpuBaseline:
    """The reference CPU path on the host cores.  Sharded by PROCESS (the reference's GJK scratch and world
    constants are file-static/global); every process loads its own copy of the library."""

    def __init__(self, workload, nproc=None, dynamics=None):
        import multiprocessing as mp

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _libs import REF_SO, Oracle, build_oracle

        build_oracle()
        self.kind = "reference" if os.path.exists(REF_SO) else "port"
        self.x, self.vpref, self.nbr, self.w = workload
        self.A, self.B, self.L, self.E = dynamics if dynamics is not None else reference_dynamics()
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
        hulls = int(sum(r[3] for r in res))
        cb_s = float(np.sum([r[4] for r in res]))
        self.stages = {"unit": "cpu-us per pair (summed over processes / pairs)",
                       "obstacle_creation": 1e6 * st[0] / max(pairs, 1), "half_plane_creation": 1e6 * st[1] / max(pairs, 1),
                       "new_velocity": 1e6 * st[2] / max(pairs, 1), "hull_branch_share_of_half_plane": st[3] / max(st[1], 1e-30),
                       "hulls": hulls,
                       "hull_ms_each": 1e3 * st[3] / max(1, hulls),
                       # the reference asks qconvex four times per hull (MAIN:769-770, 814-815); every call parses
                       # pointList.txt, computes the hull and writes its answer, and convexHullFiles writes / parses the files
                       "hull_geometry_ms_each": 1e3 * cb_s / max(1, hulls),
                       "text_roundtrip_ms_each": 1e3 * (st[3] - cb_s) / max(1, hulls),
                       "note": "hull = in-process qhull-order quickhull behind the reference's own text round trip (4 qconvex requests "
                               "per hull), no process spawn; text_roundtrip = convexHull time minus the geometry"}
        return pairs, dt, len(agents)

    def close(self):
        self.pool.terminate()


def measure_spawn_ms(n_hulls=2):
    """The reference's real hull path: convexHull spawning `qconvex` four times per hull (MAIN:769-770, 814-815), with a
    scipy/qhull-backed executable on PATH (oracle/qconvex_scipy; qhull 2012.1 ships only as a Windows binary).
    Returns ms per hull, or None where oracle/_ref is absent."""
    import shutil
    import tempfile

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _libs import REF_SO, Reference

    if not os.path.exists(REF_SO):
        return None
    d = tempfile.mkdtemp(prefix="lqro_qconvex_")
    os.symlink(os.path.join(ROOT, "oracle", "qconvex_scipy"), os.path.join(d, "qconvex"))
    old = os.environ.get("PATH", "")
    os.environ["PATH"] = d + os.pathsep + os.path.dirname(sys.executable) + os.pathsep + old
    try:
        ref = Reference()
        rng = np.random.default_rng(7)
        ref.lib.ref_set_system_passthrough(1)
        t = []
        for h in range(n_hulls + 1):
            pts = rng.normal(size=(550, 3)) * [3.0, 2.0, 1.0]   # the size of an inside cloud of the bench swarm
            t0 = time.perf_counter()
            ref.convex_hull(rng.normal(size=3) * 0.2, pts)
            t.append(time.perf_counter() - t0)
        ref.lib.ref_set_system_passthrough(0)
        return 1e3 * float(np.mean(t[1:]))   # the first hull's two stale-input requests fail early (no pointList.txt yet)
    except Exception:
        return None
    finally:
        os.environ["PATH"] = old
        shutil.rmtree(d, ignore_errors=True)


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
def sim_update_arm(ctx, x, L, E, Lh, Eh, stream, args):
    """The simulator-form update (k_sim_flow, k_sim_noise, k_sim_filter) over every agent of the swarm: estimates = the swarm's states, identity attitudes, P = 1e-9 I, synthetic
    standard normals (numpy; the kernel's cost does not depend on their values), gains from the golden vectors.  Timed
    with CUDA events around each launch (lqro_sim_last_ms); beside it the reference's own update loop on one host core
    (oracle/_ref, 32 quads) when it is available."""
    import torch
    from lqr_obstacles_b200.api import sim_default_model
    n = len(x)
    ctx.sim_set_model(sim_default_model(L=L, E=E, Lh=Lh, Eh=Eh))
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
    arm["timestep"] = {"ms_per_step": float(np.mean(ts[2:])), "api": "lqro_timestep_device (MAIN:1373-1417 for every agent: k_hull polling + k_pairs, k_hull, "
                       "k_new_velocity, k_sim_control, k_sim_flow, k_sim_noise, k_sim_filter; state resident)",
                       "gpu_launches_per_step": 8, "finite": bool(torch.isfinite(dvg).all().item())}
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


def u64_sum(a):
    """order-independent exact checksum: wrap-around sum of the raw 64-bit patterns (identical for any partition of the agents)"""
    return int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64))


def status_hash(status, first_pair):
    """wrap-around sum over pairs of (status & 0xff) * odd multiplier of the GLOBAL pair index"""
    st = (np.ascontiguousarray(status).ravel() & 0xFF).astype(np.uint64)
    idx = np.arange(first_pair, first_pair + st.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int((st * (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))).sum(dtype=np.uint64))


def oracle_sample_check(ctx, x, vpref, nbr, agents_global, lo, newV_local, res):
    """UNTIMED checker: the listed agents' whole rows (status, half-planes, new velocity) against the CPU oracle, bit for bit.
    The oracle is test infrastructure; it is never the thing measured."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _libs import Oracle

    O = Oracle()
    NF, T, sph = ctx.tables()
    rows = nbr[agents_global]
    others = np.unique(rows[rows >= 0])
    ids = np.concatenate([agents_global, others])
    remap = {int(g): len(agents_global) + t for t, g in enumerate(others)}
    for t, g in enumerate(agents_global):
        remap[int(g)] = t
    rr = np.array([[remap[int(j)] if j >= 0 else -1 for j in r] for r in rows], dtype=np.int32)
    full = np.full((len(ids), nbr.shape[1]), -1, dtype=np.int32)
    full[:len(agents_global)] = rr
    vs = np.zeros((len(ids), 3))
    vs[:len(agents_global)] = vpref[agents_global]
    oV, opl, ost, _ = O.step(NF, T, sph, np.ascontiguousarray(x[ids]), vs, full, nthreads=min(8, len(os.sched_getaffinity(0))))
    n = len(agents_global)
    loc = agents_global - lo
    st = res["status"][loc] & 0xFF
    has = (ost[:n] & 0xFF) != 0
    ok = (np.array_equal(st, ost[:n] & 0xFF) and np.array_equal(res["planes"][loc][has], opl[:n][has])
          and np.array_equal(newV_local[loc], oV[:n]))
    return bool(ok), int(has.sum()), int(((ost[:n] & 0xFF) == 2).sum())


def riccati_arm(ctx, n_batch=4096):
    """Batched Riccati solve (north_star subsystem 1; controlMatrices MAIN:512-543): n_batch HETEROGENEOUS problems -- A, B from
    lqro_linearize_discretize about perturbed operating points -- through lqro_solve_gains; kernel time from CUDA events around
    the launch (lqro_riccati_last_ms), wall time of the host-buffer call beside it, gains of a few problems against the oracle."""
    from lqr_obstacles_b200.api import sim_default_model

    m = sim_default_model()
    u0 = float(m.u_goal[0])
    rng = np.random.default_rng(4096)
    xs = rng.normal(size=(n_batch, 16)) * 0.05
    xs[:, 12:] = u0 * (1.0 + 0.05 * rng.normal(size=(n_batch, 4)))
    Rs = np.tile(np.eye(3).ravel(), (n_batch, 1))
    us = u0 * (1.0 + 0.05 * rng.normal(size=(n_batch, 4)))
    A, B, _ = ctx.linearize_discretize(xs, Rs, us)
    ctx.solve_gains(A[:64], B[:64])  # warm
    t0 = time.perf_counter()
    L, E, Lh, Eh = ctx.solve_gains(A, B)
    wall = (time.perf_counter() - t0) * 1e3
    k_ms = ctx.riccati_last_ms()
    one0 = time.perf_counter()
    ctx.solve_gains(A[:1], B[:1])
    one_wall = (time.perf_counter() - one0) * 1e3
    one_ms = ctx.riccati_last_ms()
    err = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _libs import Oracle

        O = Oracle()
        err = 0.0
        for i in rng.choice(n_batch, 6, replace=False):
            oL, oE, oLh, oEh = O.riccati(A[i], B[i])
            for got, want in ((L[i], oL), (E[i], oE), (Lh[i], oLh), (Eh[i], oEh)):
                err = max(err, float(np.abs(got - want).max() / np.abs(want).max()))
    except Exception:
        pass
    return {"problems": n_batch, "kernel_ms": k_ms, "problems_per_s": n_batch / (k_ms * 1e-3), "wall_ms_host_buffers": wall,
            "batch_of_one": {"kernel_ms": one_ms, "wall_ms": one_wall, "problems_per_s": 1.0 / (one_ms * 1e-3)},
            "iterations": 300, "finite": bool(np.isfinite(L).all() and np.isfinite(Lh).all()),
            "max_rel_err_vs_oracle_6_problems": err, "tolerance": 1e-9,
            "api": "lqro_linearize_discretize + lqro_solve_gains (k_riccati, warp per problem); heterogeneous operating points"}


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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # long before the timed region: nvidia-smi takes a few hundred ms to come up
    x, vpref, nbr, w = make_workload(args.workload)
    N, K, H, P = len(x), w["K"], w["H"], w["P"]
    assert N % world == 0
    NL = N // world
    lo = rank * NL
    ctx = Lqro(device=local_rank, n_agents=N, local_begin=lo, n_local=NL, max_neighbors=K, horizon=H, n_samples=P)
    # the dynamics come from the repo's own controlMatrices on the device (not from the golden file)
    A, B, L, E, Lh, Eh, gains_dev = device_dynamics(ctx)
    ctx.set_dynamics(A, B, L, E)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(Lqro.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    p2p_gather = bool(world > 1 and ctx.comm_p2p_active())
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

    def one_step(ev_ag=None):
        if ev_ag:
            ev_ag[0].record(stream)
        ctx.allgather_state()  # no-op on one GPU
        if ev_ag:
            ev_ag[1].record(stream)
        ctx.step_device(None, d_vpref.data_ptr(), d_nbr.data_ptr(), d_newV.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            one_step()
    barrier()
    mk = lambda: torch.cuda.Event(enable_timing=True)
    ev = [(mk(), mk()) for _ in range(args.steps)]
    ev_ag = [(mk(), mk()) for _ in range(args.steps)]
    kern_ms = np.zeros(4)
    detail_ms = np.zeros(3)
    barrier()
    sampler.begin()
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            if world > 1:
                dist.barrier()  # line the ranks up before every timed step: host-side skew is not step time
            flush.zero_()  # L2 flush, outside the per-step events
            ev[s][0].record(stream)
            one_step(ev_ag[s])
            ev[s][1].record(stream)
            stream.synchronize()
            t = ctx.last_step_ms()
            kern_ms += np.array([t["pairs"], t["hull"], t["lp"], t["total"]])
            dt = ctx.last_step_detail_ms()
            detail_ms += np.array([dt["setup"], dt["pairs"], dt["polling_hull"]])
            overlap_sms, first_slots = dt["overlap_sms"], dt["first_slots"]
    barrier()
    sampler.end()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    ag_ms = sum(a.elapsed_time(b) for a, b in ev_ag) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    kern_ms /= args.steps
    detail_ms /= args.steps
    counters = ctx.counters()
    res_dev = ctx.pair_results()
    newV_dev = d_newV.cpu().numpy()

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
    e2e_launches = ctx.counters()["launches"]
    nbr_streamed = ctx.last_step_nbr_streamed()

    # ---- checksums that do not depend on the number of ranks (wrap-around 64-bit sums, all-reduced) and an UNTIMED
    # oracle sample of this rank's agents
    hout_np = hout.numpy()
    same_arms = bool(np.array_equal(hout_np, newV_dev))
    sums = torch.tensor([np.int64(np.uint64(u64_sum(hout_np)).astype(np.int64)),
                         np.int64(np.uint64(status_hash(res_dev["status"], lo * K)).astype(np.int64))], dtype=torch.int64, device="cuda")
    fsum = torch.tensor([float(hout_np.sum()), float(counters["inside"]), float(counters["planes"]), float(counters["degenerate"]),
                         float(counters["hull_overflow"])], dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(1000 + rank)
    sample = np.sort(lo + rng.choice(NL, size=min(args.parity_agents, NL), replace=False))
    inside_owner = np.unique(np.argwhere((res_dev["status"] & 0xFF) == 2)[:, 0])[:8] + lo   # a few agents that own a hull plane
    sample = np.unique(np.concatenate([sample, inside_owner])).astype(np.int64)
    p_ok, p_planes, p_inside = oracle_sample_check(ctx, x, vpref, nbr, sample, lo, newV_dev, res_dev)
    pstat = torch.tensor([1 if (p_ok and same_arms) else 0, len(sample), p_planes, p_inside], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        pmin = pstat[:1].clone()
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(pstat, op=dist.ReduceOp.SUM)
        pstat[0] = pmin[0]
    sums, fsum, pstat = sums.cpu().numpy(), fsum.cpu().numpy(), pstat.cpu().numpy()

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
                   "h2d_bytes_per_step": int(x_loc.nbytes + vp_loc.nbytes), "gpu_launches_per_step": 8,
                   "neighbour_table_identical": bool(np.array_equal(ctx2.neighbors(), nbr_loc)),
                   "newV_identical": bool(np.array_equal(hout2.numpy(), hout_np)), "api": "lqro_step_sharded(nbr=NULL) with LQRO_FLAG_DEVICE_KNN"}
        ctx2.close()

    # ---- the second half of the reference's time step (MAIN:1408-1417: LQR feedback, propagate, Kalman filter) for the same
    # swarm on the device (lqro_sim_update_device) and the batched Riccati solve: informative, not part of `value`
    sim_arm = ric_arm = None
    if world == 1:
        sim_arm = sim_update_arm(ctx, x, L, E, Lh, Eh, stream, args)
        ric_arm = riccati_arm(ctx)

    t_dev = torch.tensor([dev_ms, e2e_ms, e2e_wall_ms, kern_ms[0], kern_ms[1], kern_ms[2], ag_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_wall_ms, pair_ms, hull_ms, lp_ms, ag_ms = [float(v) for v in t_dev.cpu()]
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
        # the dominant launch: k_pairs, one launch over every pair of the rank; on a large swarm it runs on the SMs the polling
        # hull launch does not use -- its duration, the FP64 peak of ITS SMs
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        if overlap_sms > 0 and detail_ms[2] > 0:
            pairs_per_launch = NL * K
            pair_s = float(detail_ms[1]) * 1e-3
            sm_frac = (sm_count - overlap_sms) / sm_count
        else:
            pairs_per_launch = NL * K
            pair_s = pair_ms * 1e-3
            sm_frac = 1.0
        achieved_tf = pairs_per_launch * f_alg(H, P) / pair_s / 1e12
        fp64_peak_full = fp64_peak
        fp64_peak = fp64_peak * sm_frac
        # DRAM traffic and executed-FP64 counts of k_pairs come from the committed ncu captures of this workload at this rank
        # count (profiles/rNN_k_pairs_<workload>[_nG]_ncu.json, written by tools/ from the .ncu-rep); never measured under ncu here
        ncu = None
        import glob
        caps = [c for c in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_pairs_%s*_ncu.json" % args.workload)))]
        for cfile in reversed(caps):
            try:
                cand = json.load(open(cfile))
            except Exception:
                continue
            if cand.get("pairs_per_launch") == pairs_per_launch:
                ncu = cand
                ncu["file"] = os.path.relpath(cfile, ROOT)
                break
        roofline = {
            "bound": "fp64", "kernel": "k_pairs", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak,
            "traffic": (ncu["traffic_bytes"] if ncu else None),
            "traffic_note": (None if ncu else "no committed ncu capture of k_pairs at %d pairs per launch (captures are taken at 1 GPU)" % pairs_per_launch),
            "algorithmic_bytes_per_launch": pairs_per_launch * b_alg(K),
            "executed": ({"source": ncu["file"], "fp64_flop_per_launch": ncu["executed_fp64_flop"],
                          "fp64_pipe_pct": ncu["fp64_pipe_pct"], "issue_active_pct": ncu["issue_active_pct"],
                          "note": "exact pruning skips horizon steps, so executed flop < algorithmic flop (F_alg counts every step)"}
                         if ncu else None),
            "algorithmic_flop_per_pair": f_alg(H, P), "pairs_per_launch": pairs_per_launch, "kernel_ms": pair_s * 1e3,
            "launch": ("k_pairs, one launch over every pair of the rank, on %d of %d SMs (the other %d run the polling hull launch, a "
                       "concurrent consumer of the hull queue); peak = measured whole-chip FP64 peak %.2f TFLOP/s x %d/%d"
                       % (sm_count - overlap_sms, sm_count, overlap_sms, fp64_peak_full, sm_count - overlap_sms, sm_count)
                       if sm_frac < 1.0 else "k_pairs, one launch over every pair of the rank"),
            "overlap_ms": {"k_pairs": float(detail_ms[1]), "polling_hull_launch": float(detail_ms[2]),
                           "sms_of_the_polling_hull_launch": int(overlap_sms), "first_chunk_slots": int(first_slots)},
            "peak_source": "FP64 CUDA-core pipe measured in this run (DFMA chains, lqro_measure_fp64_peak; ncu capture of that kernel: "
                           "profiles/r2_k_peak_fp64_ncu.txt); MEASURED_PEAKS.json has no FP64 figure",
            "hbm": {"achieved": pairs_per_launch * b_alg(K) / pair_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": pairs_per_launch * b_alg(K) / pair_s / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
            "fp32_peak_tflops": fp32_peak,
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu:
            cpu = CpuBaseline((x, vpref, nbr, w), dynamics=(A, B, L, E))
            cpu.run(cpu.cores)  # warm the pool
            p, d, na = cpu.run(args.cpu_agents)
            cpu.close()
            cpu_baseline = {"value": p / d, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "stages": getattr(cpu, "stages", None),
                            "text_roundtrip_ms": (cpu.stages or {}).get("text_roundtrip_ms_each"),
                            "spawn_ms": measure_spawn_ms(),
                            "spawn_note": "ms per hull when convexHull really spawns `qconvex` four times (MAIN:769-770, 814-815; "
                                          "scipy/qhull-backed executable, oracle/qconvex_scipy); not part of `value`",
                            "sample": "%d randomly chosen agents (x K=%d pairs + LP) of the same swarm, process-sharded, %.1f s" % (na, K, d)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d agents, K=%d, H=%d, P=%d, jittered 4 m lattice, seed 20261018" % (args.workload, N, K, H, P),
                       "parallelism": "agents sharded x%d, all-gather of the 128-B state per step: %s" % (
                           world, "every rank pushes its block into the peers' state buffers over NVLink peer memory (csrc/lqro_p2p.cuh)"
                           if p2p_gather else "ncclAllGather"),
                       "l2": "256 MiB buffer written between timed steps (outside the per-step events)",
                       "pairs_per_step": total_pairs,
                       "dynamics": "controlMatrices on the device (lqro_linearize_discretize + lqro_solve_gains); max relative "
                                   "deviation of A, B, L, E from the reference's: %.2e" % gains_dev},
            "clocks": clocks,
            "stages_ms": {"allgather": ag_ms, "pairs": pair_ms, "hull": hull_ms, "lp": lp_ms,
                          "note": "CUDA events on the launching stream, mean over the timed steps, max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * (x_loc.nbytes + vp_loc.nbytes + nbr_loc.nbytes)),
                    "d2h_bytes_per_step": int(world * NL * 24), "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps,
                    "api": "lqro_step_sharded (host pinned buffers: state, preferred velocities and neighbour table up, new velocities down, every step)",
                    "neighbour_table_upload": ("next to the pair kernel: rows in two blocks on a copy stream, the kernel walks the first "
                                               "eighth of the agents first and waits for the rows of a batch (lqro_last_step_nbr_streamed)"
                                               if nbr_streamed else "in front of the pair kernel"),
                    "gpu_launches_per_step": e2e_launches, "device_knn": knn_arm},
            "checksum": {"newV_u64": "%016x" % (int(sums[0]) & 0xFFFFFFFFFFFFFFFF), "status_u64": "%016x" % (int(sums[1]) & 0xFFFFFFFFFFFFFFFF),
                         "newV_float_sum": float(fsum[0]), "inside_pairs": int(fsum[1]), "planes": int(fsum[2]),
                         "degenerate": int(fsum[3]), "hull_overflow": int(fsum[4]),
                         "note": "wrap-around 64-bit sums over every rank's new velocities / pair statuses: identical for 1, 2, 4, 8 GPUs"},
            "parity_sample_ok": bool(pstat[0] == 1),
            "parity_sample": {"agents": int(pstat[1]), "planes_compared": int(pstat[2]), "inside_pairs_compared": int(pstat[3]),
                              "what": "untimed: whole rows (status, half-planes, newV) of random agents of every rank + agents owning a "
                                      "hull plane vs the CPU oracle, bit for bit; device-resident and host-buffer arms equal"},
            # per rank and step: k_hull polling, k_pairs, k_hull wide, k_new_velocity; with the peer-memory all-gather also
            # k_p2p_gather + k_p2p_done (lqro_p2p.cuh) -- with the fallback the all-gather kernel is NCCL's and not counted
            "gpu_launches": (counters["launches"] + (2 if p2p_gather else 0)) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "counters": counters,
            "riccati": ric_arm,
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
    ap.add_argument("--parity-agents", type=int, default=96, help="agents per rank in the untimed oracle sample")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
