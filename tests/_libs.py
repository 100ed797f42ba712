"""ctypes bindings used by the tests (and by bench.py's cpu_baseline leg) for the two CHECKERS:

* ``oracle/liblqro_oracle.so``  -- our CPU restatement (always buildable: plain C++);
* ``oracle/_ref/liblqro_ref.so`` -- the reference's own sources compiled from /root/reference
  (present in the build container and, as a prebuilt file, on the GPU box; never in git).

Nothing here is imported by the product package.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblqro_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "liblqro_ref.so")
REFERENCE_TREE = "/root/reference/QuadrotorHoverController"

c_int, c_double, c_void_p = ctypes.c_int, ctypes.c_double, ctypes.c_void_p


def P(a):
    """pointer to a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_void_p)


def build_oracle(force=False):
    """Compile oracle/ (and oracle/_ref when /root/reference is present).  Building the checker is
    not using it."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.isdir(REFERENCE_TREE) and not os.path.exists(REF_SO)
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "all"], stdout=subprocess.DEVNULL)


class PairOut(ctypes.Structure):
    _fields_ = [
        ("status", c_int), ("n_reach", c_int), ("best_facet", c_int), ("n_facets", c_int), ("degenerate", c_int),
        ("distance", c_double), ("normal", c_double * 3), ("plane", ctypes.c_float * 6), ("gjk_stats", c_int * 6),
    ]


class HullResult(ctypes.Structure):
    _fields_ = [
        ("nfacets", c_int), ("nvertices", c_int), ("ninserted", c_int), ("degenerate", c_int),
        ("planes", ctypes.POINTER(c_double)), ("verts", ctypes.POINTER(c_int)), ("insert_order", ctypes.POINTER(c_int)),
    ]


class Oracle:
    """Thin wrapper over liblqro_oracle.so (the CPU restatement)."""

    def __init__(self):
        build_oracle()
        self.lib = lib = ctypes.CDLL(ORACLE_SO)
        lib.lqro_o_gjk.restype = c_double
        lib.lqro_round6.restype = c_double
        lib.lqro_round6.argtypes = [c_double]
        lib.lqro_o_spheres.argtypes = [c_int, c_double, c_double, c_double, c_void_p]
        lib.lqro_o_horizon.argtypes = [c_void_p] * 4 + [c_int] + [c_void_p] * 3
        lib.lqro_o_candidates.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.lqro_o_reachable.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_double, c_void_p]
        lib.lqro_o_gjk.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        lib.lqro_o_hull_escape.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.lqro_o_pair.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_int,
                                    c_void_p, ctypes.POINTER(PairOut)]
        lib.lqro_o_new_velocity.argtypes = [c_void_p, c_int, c_void_p, c_double, c_void_p]
        lib.lqro_o_riccati.argtypes = [c_void_p, c_void_p, c_double, c_double, c_double, c_int] + [c_void_p] * 4
        lib.lqro_o_step.argtypes = [c_int] * 4 + [c_void_p] * 6 + [c_double, c_double, c_int, c_int] + [c_void_p] * 4
        lib.lqro_o_agent_update.argtypes = [c_int, c_void_p, c_void_p, c_void_p, ctypes.c_float]
        lib.lqro_hull_qorder.argtypes = [c_void_p, c_int, ctypes.POINTER(HullResult)]

    def spheres(self, P_=50, rx=0.26, ry=0.26, rz=0.75):
        out = np.zeros((P_, 3))
        self.lib.lqro_o_spheres(P_, rx, ry, rz, P(out))
        return out

    def horizon(self, A, B, L, E, H):
        NF, T, CG = np.zeros((H, 3, 16)), np.zeros((H, 3, 3)), np.zeros((H, 3, 3))
        self.lib.lqro_o_horizon(P(A), P(B), P(L), P(E), H, P(NF), P(T), P(CG))
        return NF, T, CG

    def candidates(self, NF, T, sph, xi, xj):
        H, P_ = NF.shape[0], sph.shape[0]
        q = np.zeros((H * P_, 3))
        self.lib.lqro_o_candidates(P(NF), P(T), H, P_, P(sph), P(xi), P(xj), P(q))
        return q

    def reachable(self, q, xi, xj, max_speed=30.0):
        idx = np.zeros(len(q), dtype=np.int32)
        n = self.lib.lqro_o_reachable(P(q), len(q), P(xi), P(xj), max_speed, P(idx))
        return idx[:n].copy()

    def gjk(self, vrel, pts):
        w1, w2, st = np.zeros(3), np.zeros(3), np.zeros(6, dtype=np.int32)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        d2 = self.lib.lqro_o_gjk(P(np.ascontiguousarray(vrel, dtype=np.float64)), P(pts), len(pts), P(w1), P(w2), P(st))
        return d2, w1, w2, st

    def hull(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        r = HullResult()
        rc = self.lib.lqro_hull_qorder(P(pts), len(pts), ctypes.byref(r))
        if rc != 0:
            self.lib.lqro_hull_free(ctypes.byref(r))
            raise RuntimeError("hull failed rc=%d" % rc)
        planes = np.ctypeslib.as_array(r.planes, (r.nfacets, 4)).copy()
        verts = np.ctypeslib.as_array(r.verts, (r.nfacets, 3)).copy()
        order = np.ctypeslib.as_array(r.insert_order, (max(r.ninserted, 1),)).copy()[: r.ninserted]
        info = dict(nvertices=r.nvertices, ninserted=r.ninserted, degenerate=r.degenerate)
        self.lib.lqro_hull_free(ctypes.byref(r))
        return planes, verts, order, info

    def round6(self, a):
        a = np.asarray(a, dtype=np.float64)
        return np.array([self.lib.lqro_round6(float(v)) for v in a.ravel()]).reshape(a.shape)

    def hull_escape(self, pts, vrel, fix_facet0=1, prev_normal=None):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        dist = c_double()
        nrm = np.zeros(3) if prev_normal is None else np.array(prev_normal, dtype=np.float64)
        best, deg = c_int(-1), c_int(0)
        nf = self.lib.lqro_o_hull_escape(P(pts), len(pts), P(np.ascontiguousarray(vrel, dtype=np.float64)), fix_facet0,
                                         ctypes.byref(dist), P(nrm), ctypes.byref(best), ctypes.byref(deg))
        return nf, dist.value, nrm, best.value, deg.value

    def pair(self, NF, T, sph, xi, xj, max_speed=30.0, fix_facet0=1, prev_normal=None):
        out = PairOut()
        pn = None if prev_normal is None else np.array(prev_normal, dtype=np.float64)
        self.lib.lqro_o_pair(P(NF), P(T), NF.shape[0], sph.shape[0], P(sph), P(xi), P(xj), max_speed, fix_facet0,
                             P(pn), ctypes.byref(out))
        return dict(status=out.status, n_reach=out.n_reach, best_facet=out.best_facet, n_facets=out.n_facets,
                    degenerate=out.degenerate, distance=out.distance, normal=np.array(out.normal[:]),
                    plane=np.array(out.plane[:], dtype=np.float32), gjk_stats=np.array(out.gjk_stats[:]))

    def new_velocity(self, planes, goal, radius=100.0):
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(3)
        self.lib.lqro_o_new_velocity(P(planes), len(planes), P(np.ascontiguousarray(goal, dtype=np.float64)), radius, P(out))
        return out

    def riccati(self, A, B, qv=100.0, qp=0.1, r=5.0, iters=300):
        L, E, Lh, Eh = np.zeros((4, 16)), np.zeros((4, 3)), np.zeros((3, 16)), np.zeros((3, 3))
        self.lib.lqro_o_riccati(P(A), P(B), qv, qp, r, iters, P(L), P(E), P(Lh), P(Eh))
        return L, E, Lh, Eh

    def step(self, NF, T, sph, x, vpref, nbr, max_speed=30.0, lp_radius=100.0, fix_facet0=1, nthreads=1):
        N, K = nbr.shape
        newV = np.zeros((N, 3))
        planes = np.zeros((N, K, 6), dtype=np.float32)
        status = np.zeros((N, K), dtype=np.int32)
        nplanes = np.zeros(N, dtype=np.int32)
        self.lib.lqro_o_step(N, K, NF.shape[0], sph.shape[0], P(NF), P(T), P(sph), P(x), P(vpref), P(nbr), max_speed,
                             lp_radius, fix_facet0, nthreads, P(newV), P(planes), P(status), P(nplanes))
        return newV, planes, status, nplanes

    def knn(self, x, K, begin=0, n_local=None, nthreads=1):
        """AGENT:153-174 neighbour lists by brute force (lqro_o_knn)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n_local = len(x) - begin if n_local is None else n_local
        nbr = np.zeros((n_local, K), dtype=np.int32)
        self.lib.lqro_o_knn(len(x), P(x), begin, n_local, K, nthreads, P(nbr))
        return nbr


class Reference:
    """Wrapper over oracle/_ref/liblqro_ref.so: the reference's own code (see oracle/ref_harness.cpp).
    NOT thread-safe and not re-entrant (GJK scratch is file-static, GJK.c:163-164; world constants
    are globals, MAIN:40-76; the hull files have fixed names, MAIN:764,774,790): one per process."""

    def __init__(self, workdir=None):
        build_oracle()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = lib = ctypes.CDLL(REF_SO)
        lib.ref_nominal_input.restype = c_double
        lib.ref_dt.restype = c_double
        lib.ref_sim_create.restype = c_void_p
        lib.ref_sim_create.argtypes = [c_int, c_void_p, c_void_p, ctypes.c_uint]
        lib.ref_sim_destroy.argtypes = [c_void_p]
        lib.ref_sim_get.argtypes = [c_void_p] * 7
        lib.ref_sim_get_dynamics.argtypes = [c_void_p] * 7
        lib.ref_sim_set_x.argtypes = [c_void_p, c_void_p]
        lib.ref_sim_set_vgoal.argtypes = [c_void_p, c_void_p]
        lib.ref_sim_avoid.argtypes = [c_void_p, c_int, c_double] + [c_void_p] * 5
        lib.ref_sim_set_newv.argtypes = [c_void_p, c_void_p]
        lib.ref_sim_update.argtypes = [c_void_p]
        lib.ref_pair.argtypes = [c_void_p] * 6 + [c_int, c_double] + [c_void_p] * 2 + [ctypes.POINTER(c_int)] * 2 + [
            ctypes.POINTER(c_double)] + [c_void_p] * 3
        lib.ref_run_gjk.argtypes = [c_void_p, c_void_p, c_int, ctypes.POINTER(c_double), c_void_p, c_void_p]
        lib.ref_convex_hull.argtypes = [c_void_p, c_void_p, c_int, ctypes.POINTER(c_double), c_void_p]
        lib.ref_calculate_new_v.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
        lib.ref_control_matrices.argtypes = [c_void_p] * 10
        lib.ref_find_fg.argtypes = [c_void_p] * 6
        lib.ref_set_weights.argtypes = [c_double] * 3
        lib.ref_sim_facet0_events.restype = ctypes.c_long
        lib.ref_sim_set_fix_facet0.argtypes = [c_int]
        lib.ref_step_agents.restype = ctypes.c_long
        lib.ref_step_agents.argtypes = [c_void_p] * 8 + [c_int, c_int, c_int, c_double, c_void_p, c_void_p]
        lib.ref_get_seed.restype = ctypes.c_uint
        lib.ref_normals.argtypes = [c_int, c_void_p]
        lib.ref_sim_set_state.argtypes = [c_void_p] * 6
        lib.ref_sim_get_goals.argtypes = [c_void_p] * 3
        lib.ref_sim_model.argtypes = [c_void_p, c_void_p]
        for name in ("ref_exp16", "ref_exp3"):
            getattr(lib, name).argtypes = [c_void_p, c_void_p]
        lib.ref_jacobi16.argtypes = [c_void_p] * 3
        lib.ref_f.argtypes = [c_void_p] * 4
        lib.ref_linearize_discretize.argtypes = [c_void_p] * 6
        if hasattr(lib, "ref_hull_callback_seconds"):
            lib.ref_hull_callback_seconds.restype = c_double
            lib.ref_hull_callback_seconds.argtypes = [c_int]
        self._tmp = workdir or tempfile.mkdtemp(prefix="lqro_ref_")
        lib.ref_chdir(self._tmp.encode())
        lib.ref_init()
        self.H_max = lib.ref_obstacle_steps()
        self.num_points = lib.ref_num_points()
        self._cb = None

    def nominal_input(self):
        return self.lib.ref_nominal_input()

    def control_matrices(self):
        u = np.full(4, self.nominal_input())
        xs = np.zeros(16)
        xs[12:] = u
        A, B, c = np.zeros((16, 16)), np.zeros((16, 4)), np.zeros(16)
        L, E, l = np.zeros((4, 16)), np.zeros((4, 3)), np.zeros(4)
        Lh, Eh = np.zeros((3, 16)), np.zeros((3, 3))
        self.lib.ref_control_matrices(P(u), P(xs), P(A), P(B), P(c), P(L), P(E), P(l), P(Lh), P(Eh))
        return dict(A=A, B=B, c=c, L=L, E=E, l=l, Lh=Lh, Eh=Eh)

    def find_fg(self, A, B, L, E, F, G):
        F, G = F.copy(), G.copy()
        self.lib.ref_find_fg(P(A), P(B), P(L), P(E), P(F), P(G))
        return F, G

    def spheres(self):
        out = np.zeros((self.num_points, 3))
        n = self.lib.ref_create_spheres(P(out))
        assert n == self.num_points
        return out

    def pair(self, A, B, L, E, xi, xj, H=30, max_speed=30.0, prev_normal=None):
        n = H * self.num_points
        ell = np.zeros((n, 3))
        ridx = np.zeros(n, dtype=np.int32)
        nre, ins, dist = c_int(), c_int(), c_double()
        nrm = np.zeros(3) if prev_normal is None else np.array(prev_normal, dtype=np.float64)
        plane = np.zeros(6, dtype=np.float32)
        st = np.zeros(6, dtype=np.int32)
        status = self.lib.ref_pair(P(A), P(B), P(L), P(E), P(xi), P(xj), H, max_speed, P(ell), P(ridx), ctypes.byref(nre),
                                   ctypes.byref(ins), ctypes.byref(dist), P(nrm), P(plane), P(st))
        return dict(status=status, n_reach=nre.value, inside=ins.value, distance=dist.value, normal=nrm, plane=plane,
                    gjk_stats=st, ellipsoids=ell, reach_idx=ridx[: nre.value].copy())

    def run_gjk(self, vrel, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        d = c_double()
        nrm = np.zeros(3)
        st = np.zeros(6, dtype=np.int32)
        self.lib.ref_run_gjk(P(np.ascontiguousarray(vrel, dtype=np.float64)), P(pts), len(pts), ctypes.byref(d), P(nrm), P(st))
        return d.value, nrm, st

    def convex_hull(self, vrel, pts, prev_normal=None):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        d = c_double()
        nrm = np.zeros(3) if prev_normal is None else np.array(prev_normal, dtype=np.float64)
        self.lib.ref_convex_hull(P(np.ascontiguousarray(vrel, dtype=np.float64)), P(pts), len(pts), ctypes.byref(d), P(nrm))
        return d.value, nrm

    def last_hull(self):
        nf, npnt = c_int(), c_int()
        self.lib.ref_last_hull(ctypes.byref(nf), ctypes.byref(npnt), None, None, 0)
        if nf.value <= 0:
            return None
        planes = np.zeros((nf.value, 4))
        verts = np.zeros((nf.value, 3), dtype=np.int32)
        self.lib.ref_last_hull(ctypes.byref(nf), ctypes.byref(npnt), P(planes), P(verts), nf.value)
        return planes, verts

    def calculate_new_v(self, planes, goal):
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(3)
        self.lib.ref_calculate_new_v(P(planes), len(planes), P(np.ascontiguousarray(goal, dtype=np.float64)), P(out))
        return out

    def step_agents(self, A, B, L, E, x, vpref, nbr_rows, agents, H=20, max_speed=30.0):
        """avoidance step of the listed agents, entirely inside the reference's code (ref_step_agents)"""
        agents = np.ascontiguousarray(agents, dtype=np.int32)
        nbr_rows = np.ascontiguousarray(nbr_rows, dtype=np.int32)
        newV = np.zeros((len(agents), 3))
        nplanes = np.zeros(len(agents), dtype=np.int32)
        npairs = self.lib.ref_step_agents(P(A), P(B), P(L), P(E), P(x), P(vpref), P(nbr_rows), P(agents), len(agents),
                                          nbr_rows.shape[1], H, max_speed, P(newV), P(nplanes))
        return npairs, newV, nplanes

    def use_scipy_hull(self, enable=True):
        """Route the qconvex stand-in to scipy.spatial.ConvexHull (qhull_r 8.0.2) instead of
        oracle/hull_cpu.cpp.  Build-container validation only."""
        if not enable:
            self.lib.ref_set_hull_callback(None)
            self._cb = None
            return
        from scipy.spatial import ConvexHull

        CB = ctypes.CFUNCTYPE(c_int, c_int, ctypes.POINTER(c_double), c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_int))

        def cb(n, pts, maxf, planes, verts):
            try:
                p = np.ctypeslib.as_array(pts, (n, 3)).copy()
                h = ConvexHull(p)
                nf = len(h.simplices)
                if nf > maxf:
                    return -2
                np.ctypeslib.as_array(planes, (nf, 4))[:] = h.equations
                np.ctypeslib.as_array(verts, (nf, 3))[:] = h.simplices
                return nf
            except Exception:
                return -1

        self._cb = CB(cb)
        self.lib.ref_set_hull_callback(self._cb)


class RefSim:
    """Run-time-sized copy of the reference main loop (oracle/ref_harness.cpp: ref_sim_*)."""

    def __init__(self, ref, starts, goals, seed=1):
        self.ref, self.lib = ref, ref.lib
        starts = np.ascontiguousarray(starts, dtype=np.float64)
        goals = np.ascontiguousarray(goals, dtype=np.float64)
        self.n = len(starts)
        self.h = self.lib.ref_sim_create(self.n, P(starts), P(goals), seed)

    def close(self):
        if self.h:
            self.lib.ref_sim_destroy(self.h)
            self.h = None

    def state(self):
        n = self.n
        x, xt, vg = np.zeros((n, 16)), np.zeros((n, 16)), np.zeros((n, 3))
        self.lib.ref_sim_get(self.h, P(x), P(xt), P(vg), None, None, None)
        return x, xt, vg

    def dynamics(self):
        A, B, L, E, Lh, Eh = (np.zeros(s) for s in [(16, 16), (16, 4), (4, 16), (4, 3), (3, 16), (3, 3)])
        self.lib.ref_sim_get_dynamics(self.h, P(A), P(B), P(L), P(E), P(Lh), P(Eh))
        return A, B, L, E, Lh, Eh

    def avoid(self, H=30, max_speed=30.0):
        n = self.n
        newV = np.zeros((n, 3))
        planes = np.zeros((n, n - 1, 6), dtype=np.float32)
        status = np.zeros((n, n - 1), dtype=np.int32)
        nplanes = np.zeros(n, dtype=np.int32)
        nreach = np.zeros((n, n - 1), dtype=np.int32)
        self.lib.ref_sim_avoid(self.h, H, max_speed, P(newV), P(planes), P(status), P(nplanes), P(nreach))
        return newV, planes, status, nplanes, nreach

    def set_newv(self, v):
        self.lib.ref_sim_set_newv(self.h, P(np.ascontiguousarray(v, dtype=np.float64)))

    def update(self):
        self.lib.ref_sim_update(self.h)

    # -- simulator-form update parity
    def full_state(self):
        """dict of x, xTrue (n,16), vGoal (n,3), Rot, RotTrue (n,9), P (n,256), pGoal, newV (n,3)"""
        n = self.n
        d = dict(x=np.zeros((n, 16)), xTrue=np.zeros((n, 16)), vGoal=np.zeros((n, 3)), Rot=np.zeros((n, 9)),
                 RotTrue=np.zeros((n, 9)), P=np.zeros((n, 256)), pGoal=np.zeros((n, 3)), newV=np.zeros((n, 3)))
        self.lib.ref_sim_get(self.h, P(d["x"]), P(d["xTrue"]), P(d["vGoal"]), P(d["Rot"]), P(d["RotTrue"]), P(d["P"]))
        self.lib.ref_sim_get_goals(self.h, P(d["pGoal"]), P(d["newV"]))
        return d

    def set_state(self, x=None, xTrue=None, Rot=None, RotTrue=None, Pcov=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x, xTrue, Rot, RotTrue, Pcov = f(x), f(xTrue), f(Rot), f(RotTrue), f(Pcov)
        self.lib.ref_sim_set_state(self.h, P(x), P(xTrue), P(Rot), P(RotTrue), P(Pcov))

    def model(self):
        """the reference's globals + quad 0's gains as the 458 doubles of lqro_sim_model"""
        out = np.zeros(458)
        k = self.lib.ref_sim_model(self.h, P(out))
        assert k == 458
        return out

    def seed(self):
        return self.lib.ref_get_seed()


def have_reference():
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_TREE)


SIMHOST_SRC = os.path.join(ROOT, "tests", "sim_host", "sim_host.cpp")
SIMHOST_SO = os.path.join(ROOT, "tests", "sim_host", "libsimhost.so")
SIM_KERNEL_SRC = os.path.join(ROOT, "lqr_obstacles_b200", "csrc", "lqro_sim.cuh")


class SimHost:
    """The kernel source csrc/lqro_sim.cuh compiled as plain C++ (one sequential lane): checks the LOGIC of the
    simulator-form update kernel where no GPU exists.  Test infrastructure; the product never links it."""

    def __init__(self):
        if (not os.path.exists(SIMHOST_SO)
                or os.path.getmtime(SIMHOST_SO) < max(os.path.getmtime(SIMHOST_SRC), os.path.getmtime(SIM_KERNEL_SRC))):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-std=c++14", "-shared",
                                   "-o", SIMHOST_SO, SIMHOST_SRC])
        self.lib = lib = ctypes.CDLL(SIMHOST_SO)
        lib.simhost_update.argtypes = [c_int] + [c_void_p] * 11
        lib.simhost_exp16.argtypes = [c_void_p] * 2
        lib.simhost_exp3.argtypes = [c_void_p] * 2
        lib.simhost_jacobi16.argtypes = [c_void_p] * 3
        lib.simhost_f.argtypes = [c_void_p] * 5
        lib.simhost_linearize.argtypes = [c_void_p] * 7
        lib.simhost_init.argtypes = [c_int] + [c_void_p] * 9
        lib.simhost_exp_with_half.argtypes = [c_void_p] * 3

    def update(self, model, st, newV, normals):
        """st: dict as RefSim.full_state(); returns the updated copy (+ vGoal, u)"""
        o = {k: np.array(st[k], dtype=np.float64, copy=True) for k in ("x", "Rot", "P", "xTrue", "RotTrue")}
        n = len(o["x"])
        o["vGoal"], o["u"] = np.zeros((n, 3)), np.zeros((n, 4))
        model = np.ascontiguousarray(model, dtype=np.float64)
        newV = np.ascontiguousarray(newV, dtype=np.float64)
        pGoal = np.ascontiguousarray(st["pGoal"], dtype=np.float64)
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        self.lib.simhost_update(n, P(model), P(o["x"]), P(o["Rot"]), P(o["P"]), P(o["xTrue"]), P(o["RotTrue"]), P(newV), P(pGoal),
                                P(normals), P(o["vGoal"]), P(o["u"]))
        return o


HULLHOST_DIR = os.path.join(ROOT, "tests", "hull_host")
HULLHOST_SO = os.path.join(HULLHOST_DIR, "libhullhost.so")
HULLHOST_SRCS = [os.path.join(HULLHOST_DIR, "hull_host.cpp"), os.path.join(HULLHOST_DIR, "warp_emul.h"),
                 os.path.join(ROOT, "lqr_obstacles_b200", "csrc", "lqro_hull_core.cuh")]


class HullHost:
    """The kernel source csrc/lqro_hull_core.cuh compiled as plain C++ on a 32-thread lockstep emulation of the warp
    primitives (tests/hull_host/warp_emul.h): checks the LOGIC of the CUDA hull where no GPU exists.  Test
    infrastructure; the product never links it."""

    def __init__(self, general=False):
        """general=True: the build that forces the general branches (owner-word ridge matching, wide visited set, shuffle-sweep
        buckets of the appends) on every cone of more than 2 / 3 / 2 facets, so that small clouds exercise them, and (two warps per
        hull) throws away every third speculative horizon of the helper warp"""
        so = HULLHOST_SO.replace(".so", "_general.so") if general else HULLHOST_SO
        extra = ["-DLQRO_HULL_PAIRMATCH_MAX=2", "-DLQRO_HULL_NARROW_WALK_MAX=3", "-DLQRO_HULL_SMALLKEY_MAX=2", "-DLQRO_HULL_SPEC_REJECT_MOD=3"] if general else []
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in HULLHOST_SRCS):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-pthread"]
                                  + extra + ["-o", so, HULLHOST_SRCS[0]])
        self.lib = lib = ctypes.CDLL(so)
        lib.hullhost_build.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
        lib.hullhost_coverage.argtypes = [c_void_p, c_int]
        lib.hullhost_round6.restype = c_double
        lib.hullhost_round6.argtypes = [c_double]

    def hull(self, pts, mode=0, cap_f=0, vrel=None, ring=0, coop=False):
        """ring > 0: size of the pending ring (a power of two below the default, to force its compaction); coop: two warps per hull
        (main + helper, as k_hull runs it)
        -> planes (nf,4), verts (nf,3), info dict, escape (distance, normal[3]) or None"""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        n = len(pts)
        maxf = 2 * n + 64
        planes, verts = np.zeros((maxf, 4)), np.zeros((maxf, 3), dtype=np.int32)
        info, esc = np.zeros(5, dtype=np.int32), np.zeros(4)
        v = None if vrel is None else np.ascontiguousarray(vrel, dtype=np.float64)
        self.lib.hullhost_build(P(pts), n, mode | (2 if coop else 0), cap_f, ring, P(planes), P(verts), maxf, P(info), P(v), P(esc))
        nf = int(info[0])
        d = dict(nfacets=nf, degenerate=int(info[1]), fail=int(info[2]), ninserted=int(info[3]), best_facet=int(info[4]))
        return planes[:nf], verts[:nf], d, (None if vrel is None else (esc[0], esc[1:4].copy()))

    def coverage(self, reset=False):
        out = np.zeros(16, dtype=np.int64)
        self.lib.hullhost_coverage(P(out), 1 if reset else 0)
        return dict(zip(("cone_gt16", "cone_gt64", "exhaustive", "old_facet_moved", "ring_compaction", "horizon_flood",
                         "multi_chunk_partition", "multi_chunk_wave", "olist_compaction", "append_moved_segment", "append_sweep_buckets",
                         "gather_cold", "spec_adopted", "spec_rejected", "spec_none", "-"), out.tolist()))

    def round6(self, a):
        a = np.asarray(a, dtype=np.float64)
        return np.array([self.lib.hullhost_round6(float(v)) for v in a.ravel()]).reshape(a.shape)
