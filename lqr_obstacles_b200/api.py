"""ctypes mirror of include/lqro.h.  Names follow the reference's functions for the path
(obstacle creation, half-plane creation, new-velocity calculation, agent update)."""
import ctypes
import os

import numpy as np

from . import build as _build

c_int, c_double, c_float, c_void_p = ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_void_p

HULL_TEXT6, HULL_FP64 = 0, 1
PAIR_NONE, PAIR_OUTSIDE, PAIR_INSIDE, PAIR_DEGENERATE, PAIR_DROPPED = 0, 1, 2, 0x100, 0x200
FLAG_NO_PRUNE = 1
FLAG_FUSED_HULL = 2
FLAG_DEVICE_KNN = 4
FLAG_NO_OVERLAP = 8


class LqroError(RuntimeError):
    pass


class Config(ctypes.Structure):
    _fields_ = [
        ("device", c_int), ("n_agents", c_int), ("local_begin", c_int), ("n_local", c_int), ("max_neighbors", c_int),
        ("horizon", c_int), ("n_samples", c_int), ("xy_radius", c_double), ("z_radius", c_double), ("max_speed", c_double),
        ("lp_radius", c_double), ("hull_input", c_int), ("fix_facet0_normal", c_int), ("max_hull_pairs", c_int),
        ("hull_slots", c_int), ("flags", c_int),
    ]


_lib = None

EXPORTS = [
    "lqro_default_config", "lqro_create", "lqro_destroy", "lqro_last_error", "lqro_version", "lqro_solve_gains",
    "lqro_set_dynamics", "lqro_get_tables", "lqro_get_spheres", "lqro_step", "lqro_step_device", "lqro_synchronize", "lqro_set_stream",
    "lqro_step_sharded", "lqro_get_pair_results",
    "lqro_get_counters", "lqro_find_fg", "lqro_create_obstacle", "lqro_candidates", "lqro_reachable", "lqro_gjk", "lqro_convex_hull", "lqro_hull_escape",
    "lqro_new_velocity", "lqro_agent_update", "lqro_round6", "lqro_comm_unique_id", "lqro_comm_init", "lqro_comm_p2p_active", "lqro_device_state",
    "lqro_upload_local_state", "lqro_allgather_state", "lqro_measure_fp64_peak", "lqro_measure_fp32_peak", "lqro_last_step_ms", "lqro_last_step_detail_ms",
    "lqro_last_step_nbr_streamed",
    "lqro_debug_hull_cycles", "lqro_debug_hull_longest", "lqro_riccati_last_ms", "lqro_knn", "lqro_knn_device", "lqro_get_neighbors",
    "lqro_set_dynamics_class", "lqro_set_agent_classes",
    "lqro_sim_default_model", "lqro_sim_set_model", "lqro_sim_update", "lqro_sim_update_device", "lqro_msvc_normals",
    "lqro_sim_last_ms", "lqro_debug_sim_cycles", "lqro_timestep_device",
    "lqro_linearize_discretize", "lqro_sim_init",
]


class SimModel(ctypes.Structure):
    """lqro_sim_model: the reference's world / quadrotor globals (MAIN:40-76, 176-198), gains and noise covariances."""
    _fields_ = [
        ("dt", c_double), ("gravity", c_double), ("mass", c_double), ("moment_const", c_double), ("thrust_latency", c_double),
        ("length", c_double), ("jstep", c_double), ("inertia", c_double * 9), ("inv_inertia", c_double * 9),
        ("u_goal", c_double * 4), ("L", c_double * 64), ("E", c_double * 12), ("l", c_double * 4), ("Lh", c_double * 48),
        ("Eh", c_double * 9), ("M", c_double * 256), ("N", c_double * 36),
    ]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float64).copy()

    @classmethod
    def from_array(cls, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size * 8 == ctypes.sizeof(cls)
        return cls.from_buffer_copy(a.tobytes())


def sim_default_model(L=None, E=None, Lh=None, Eh=None, l=None):
    """setup() constants (MAIN:176-198), M = N = 1e-9 I, plus the given gains."""
    lib = load_library()
    m = SimModel()
    lib.lqro_sim_default_model(ctypes.byref(m))
    for name, val in (("L", L), ("E", E), ("Lh", Lh), ("Eh", Eh), ("l", l)):
        if val is not None:
            v = np.ascontiguousarray(val, dtype=np.float64).ravel()
            getattr(m, name)[:] = list(v)
    return m


def msvc_normals(seed, count):
    """`count` draws of the reference's normal() (MAIN:325-336) from MSVC rand() state `seed`; returns (normals, new seed)."""
    lib = load_library()
    s = ctypes.c_uint(seed)
    out = np.zeros(count)
    rc = lib.lqro_msvc_normals(ctypes.byref(s), count, _p(out))
    if rc != 0:
        raise LqroError("lqro_msvc_normals failed (%d)" % rc)
    return out, s.value


def load_library(path=None):
    """Load liblqro.so.  Raises if the CUDA library has not been built: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("LQRO_LIB") or _build.SO
    if not os.path.exists(path):
        raise LqroError("liblqro.so is missing (%s): build it with `python -m lqr_obstacles_b200.build`; "
                        "there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    lib.lqro_last_error.restype = ctypes.c_char_p
    lib.lqro_last_error.argtypes = [c_void_p]
    lib.lqro_version.restype = ctypes.c_char_p
    lib.lqro_default_config.argtypes = [ctypes.POINTER(Config)]
    lib.lqro_default_config.restype = None
    lib.lqro_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(c_void_p)]
    lib.lqro_destroy.argtypes = [c_void_p]
    lib.lqro_solve_gains.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_double, c_double, c_double, c_int] + [c_void_p] * 4
    lib.lqro_set_dynamics.argtypes = [c_void_p] * 5
    lib.lqro_get_tables.argtypes = [c_void_p] * 4
    lib.lqro_get_spheres.argtypes = [c_void_p] * 2
    lib.lqro_step.argtypes = [c_void_p] * 5
    lib.lqro_step_device.argtypes = [c_void_p] * 6
    lib.lqro_synchronize.argtypes = [c_void_p]
    lib.lqro_set_stream.argtypes = [c_void_p, c_void_p]
    lib.lqro_step_sharded.argtypes = [c_void_p] * 5
    lib.lqro_get_pair_results.argtypes = [c_void_p] * 6
    lib.lqro_get_counters.argtypes = [c_void_p, c_void_p]
    lib.lqro_candidates.argtypes = [c_void_p] * 4
    lib.lqro_find_fg.argtypes = [c_void_p] * 7
    lib.lqro_create_obstacle.argtypes = [c_void_p, c_int] + [c_void_p] * 9
    lib.lqro_reachable.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lqro_gjk.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lqro_convex_hull.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lqro_hull_escape.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lqro_new_velocity.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lqro_agent_update.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float]
    lib.lqro_round6.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    lib.lqro_comm_unique_id.argtypes = [c_void_p]
    lib.lqro_comm_init.argtypes = [c_void_p, c_int, c_int, c_void_p]
    lib.lqro_comm_p2p_active.argtypes = [c_void_p]
    lib.lqro_device_state.argtypes = [c_void_p, ctypes.POINTER(c_void_p)]
    lib.lqro_upload_local_state.argtypes = [c_void_p, c_void_p]
    lib.lqro_allgather_state.argtypes = [c_void_p]
    lib.lqro_measure_fp64_peak.argtypes = [c_void_p, ctypes.POINTER(c_double)]
    lib.lqro_measure_fp32_peak.argtypes = [c_void_p, ctypes.POINTER(c_double)]
    lib.lqro_last_step_ms.argtypes = [c_void_p, c_void_p]
    lib.lqro_last_step_detail_ms.argtypes = [c_void_p, c_void_p]
    lib.lqro_last_step_nbr_streamed.argtypes = [c_void_p]
    lib.lqro_debug_hull_cycles.argtypes = [c_void_p, c_void_p]
    lib.lqro_debug_hull_longest.argtypes = [c_void_p, c_void_p]
    lib.lqro_sim_default_model.argtypes = [ctypes.POINTER(SimModel)]
    lib.lqro_sim_default_model.restype = None
    lib.lqro_sim_set_model.argtypes = [c_void_p, ctypes.POINTER(SimModel)]
    lib.lqro_sim_update.argtypes = [c_void_p, c_int] + [c_void_p] * 10
    lib.lqro_sim_update_device.argtypes = [c_void_p, c_int] + [c_void_p] * 11
    lib.lqro_msvc_normals.argtypes = [ctypes.POINTER(ctypes.c_uint), c_int, c_void_p]
    lib.lqro_sim_last_ms.argtypes = [c_void_p, c_void_p]
    lib.lqro_debug_sim_cycles.argtypes = [c_void_p, c_void_p, c_int]
    lib.lqro_timestep_device.argtypes = [c_void_p] * 9
    lib.lqro_linearize_discretize.argtypes = [c_void_p, c_int] + [c_void_p] * 6
    lib.lqro_sim_init.argtypes = [c_void_p, c_int] + [c_void_p] * 8
    lib.lqro_riccati_last_ms.argtypes = [c_void_p, c_void_p]
    lib.lqro_knn.argtypes = [c_void_p] * 3
    lib.lqro_knn_device.argtypes = [c_void_p] * 4   # the stream is a 64-bit handle: never let ctypes marshal it as a C int
    lib.lqro_get_neighbors.argtypes = [c_void_p] * 2
    lib.lqro_set_dynamics_class.argtypes = [c_void_p, c_int] + [c_void_p] * 4
    lib.lqro_set_agent_classes.argtypes = [c_void_p] * 2
    if path == _build.SO:
        _lib = lib
    return lib


def default_config(**kw):
    lib = load_library()
    c = Config()
    lib.lqro_default_config(ctypes.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data_as(c_void_p)
    return c_void_p(int(a))  # raw address (e.g. torch tensor data_ptr())


class Lqro:
    """One context = one GPU.  Method names are the reference's entry points."""

    def __init__(self, **cfg):
        self.lib = load_library()
        self.cfg = default_config(**cfg)
        h = c_void_p()
        rc = self.lib.lqro_create(ctypes.byref(self.cfg), ctypes.byref(h))
        if rc != 0:
            raise LqroError("lqro_create failed (%d): %s" % (rc, self.lib.lqro_last_error(None).decode()))
        self.h = h
        self.N = self.cfg.n_agents
        self.NL = self.cfg.n_local if self.cfg.n_local > 0 else self.N - self.cfg.local_begin
        self.K, self.H, self.P = self.cfg.max_neighbors, self.cfg.horizon, self.cfg.n_samples

    def close(self):
        if getattr(self, "h", None):
            self.lib.lqro_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LqroError("liblqro error %d: %s" % (rc, self.lib.lqro_last_error(self.h).decode()))

    # -- (1) controlMatrices / findFG
    def solve_gains(self, A, B, qv=100.0, qp=0.1, r=5.0, iters=300):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 16, 16)
        B = np.ascontiguousarray(B, dtype=np.float64).reshape(-1, 16, 4)
        nb = len(A)
        L, E, Lh, Eh = np.zeros((nb, 4, 16)), np.zeros((nb, 4, 3)), np.zeros((nb, 3, 16)), np.zeros((nb, 3, 3))
        self._ck(self.lib.lqro_solve_gains(self.h, nb, _p(A), _p(B), qv, qp, r, iters, _p(L), _p(E), _p(Lh), _p(Eh)))
        return L, E, Lh, Eh

    def riccati_last_ms(self):
        t = np.zeros(1, dtype=np.float32)
        self._ck(self.lib.lqro_riccati_last_ms(self.h, _p(t)))
        return float(t[0])

    def set_dynamics(self, A, B, L, E):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._ck(self.lib.lqro_set_dynamics(self.h, _p(f(A)), _p(f(B)), _p(f(L)), _p(f(E))))

    def tables(self):
        NF, T, sph = np.zeros((self.H, 3, 16)), np.zeros((self.H, 3, 3)), np.zeros((self.P, 3))
        self._ck(self.lib.lqro_get_tables(self.h, _p(NF), _p(T), _p(sph)))
        return NF, T, sph

    # -- (2)-(4) the step
    def step(self, x, vpref, nbr, newV=None):
        """Host buffers (numpy or raw pinned addresses): the copies are inside the call."""
        if newV is None:
            newV = np.zeros((self.NL, 3))
        self._ck(self.lib.lqro_step(self.h, _p(x), _p(vpref), _p(nbr), _p(newV)))
        return newV

    def step_device(self, d_x, d_vpref, d_nbr, d_newV, stream=None):
        self._ck(self.lib.lqro_step_device(self.h, _p(d_x), _p(d_vpref), _p(d_nbr), _p(d_newV), stream))

    def step_sharded(self, x_local, vpref, nbr, newV=None):
        if newV is None:
            newV = np.zeros((self.NL, 3))
        self._ck(self.lib.lqro_step_sharded(self.h, _p(x_local), _p(vpref), _p(nbr), _p(newV)))
        return newV

    def set_stream(self, stream):
        self._ck(self.lib.lqro_set_stream(self.h, None if stream is None else c_void_p(int(stream))))

    def synchronize(self):
        self._ck(self.lib.lqro_synchronize(self.h))

    def pair_results(self):
        NP = self.NL * self.K
        planes = np.zeros((self.NL, self.K, 6), dtype=np.float32)
        status = np.zeros((self.NL, self.K), dtype=np.int32)
        nreach = np.zeros((self.NL, self.K), dtype=np.int32)
        dist = np.zeros((self.NL, self.K))
        nfac = np.zeros((self.NL, self.K), dtype=np.int32)
        self._ck(self.lib.lqro_get_pair_results(self.h, _p(planes), _p(status), _p(nreach), _p(dist), _p(nfac)))
        return dict(planes=planes, status=status, nreach=nreach, dist=dist, nfacets=nfac)

    # -- simulator-form agent update (MAIN:1408-1417)
    def sim_set_model(self, model):
        self._ck(self.lib.lqro_sim_set_model(self.h, ctypes.byref(model)))

    def sim_update(self, x, Rot, P, xTrue, RotTrue, newV, pGoal, normals):
        """In place on x (n,16), Rot (n,9), P (n,256), xTrue (n,16), RotTrue (n,9); returns (vGoal (n,3), u (n,4))."""
        n = len(x)
        for a in (x, Rot, P, xTrue, RotTrue):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        newV, pGoal, normals = f(newV), f(pGoal), f(normals)
        assert normals.size == 22 * n
        vGoal, u = np.zeros((n, 3)), np.zeros((n, 4))
        self._ck(self.lib.lqro_sim_update(self.h, n, _p(x), _p(Rot), _p(P), _p(xTrue), _p(RotTrue), _p(newV), _p(pGoal), _p(normals),
                                          _p(vGoal), _p(u)))
        return vGoal, u

    def sim_update_device(self, n, d_x, d_Rot, d_P, d_xTrue, d_RotTrue, d_newV, d_pGoal, d_normals, d_vGoal, d_u=None, stream=None):
        self._ck(self.lib.lqro_sim_update_device(self.h, n, _p(d_x), _p(d_Rot), _p(d_P), _p(d_xTrue), _p(d_RotTrue), _p(d_newV),
                                                 _p(d_pGoal), _p(d_normals), _p(d_vGoal), _p(d_u), stream))

    def sim_init(self, x, Rot, P, pGoal, normals16):
        """setupQuadrotors' noise part + first findVGoal (MAIN:126-129, 1355): returns xTrue (n,16), RotTrue (n,9), vGoal (n,3)"""
        f = lambda a, k: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, k)
        x, Rot, P, pGoal, normals16 = f(x, 16), f(Rot, 9), f(P, 256), f(pGoal, 3), f(normals16, 16)
        n = len(x)
        xT, RT, vG = np.zeros((n, 16)), np.zeros((n, 9)), np.zeros((n, 3))
        self._ck(self.lib.lqro_sim_init(self.h, n, _p(x), _p(Rot), _p(P), _p(pGoal), _p(normals16), _p(xT), _p(RT), _p(vG)))
        return xT, RT, vG

    def linearize_discretize(self, x, Rot, u):
        """linearizeDiscretize (MAIN:425-440) about each (x, Rot, u): returns A (n,16,16), B (n,16,4), c (n,16)"""
        f = lambda a, k: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, k)
        x, Rot, u = f(x, 16), f(Rot, 9), f(u, 4)
        n = len(x)
        A, B, c = np.zeros((n, 16, 16)), np.zeros((n, 16, 4)), np.zeros((n, 16))
        self._ck(self.lib.lqro_linearize_discretize(self.h, n, _p(x), _p(Rot), _p(u), _p(A), _p(B), _p(c)))
        return A, B, c

    def timestep_device(self, d_Rot, d_P, d_xTrue, d_RotTrue, d_pGoal, d_normals, d_vgoal, stream=None):
        """one whole time step (MAIN:1373-1417) on device-resident state; the estimates are the context's state buffer"""
        self._ck(self.lib.lqro_timestep_device(self.h, _p(d_Rot), _p(d_P), _p(d_xTrue), _p(d_RotTrue), _p(d_pGoal), _p(d_normals),
                                               _p(d_vgoal), stream))

    def sim_cycles(self, reset=True):
        c = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.lqro_debug_sim_cycles(self.h, _p(c), int(reset)))
        return c

    def sim_last_ms(self, parts=False):
        """device ms of the last simulator-form update; parts=True: (total, k_sim_flow, k_sim_noise, k_sim_filter)"""
        ms = np.zeros(4, dtype=np.float32)
        self._ck(self.lib.lqro_sim_last_ms(self.h, _p(ms)))
        return tuple(float(v) for v in ms) if parts else float(ms[0])

    def counters(self):
        c = np.zeros(6, dtype=np.int64)
        self._ck(self.lib.lqro_get_counters(self.h, _p(c)))
        return dict(pairs=int(c[0]), planes=int(c[1]), inside=int(c[2]), degenerate=int(c[3]), hull_overflow=int(c[4]),
                    launches=int(c[5]))

    def last_step_ms(self):
        t = np.zeros(4, dtype=np.float32)
        self._ck(self.lib.lqro_last_step_ms(self.h, _p(t)))
        return dict(pairs=float(t[0]), hull=float(t[1]), lp=float(t[2]), total=float(t[3]))

    def last_step_detail_ms(self):
        t = np.zeros(5, dtype=np.float32)
        self._ck(self.lib.lqro_last_step_detail_ms(self.h, _p(t)))
        return dict(setup=float(t[0]), pairs=float(t[1]), polling_hull=float(t[2]), overlap_sms=int(t[3]), first_slots=int(t[4]))

    def last_step_nbr_streamed(self):
        """True if the host neighbour table of the last host-buffer step went up next to the pair kernel (lqro.h)."""
        return bool(self.lib.lqro_last_step_nbr_streamed(self.h))

    def hull_longest(self):
        c = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.lqro_debug_hull_longest(self.h, _p(c)))
        return c

    def hull_cycles(self):
        c = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.lqro_debug_hull_cycles(self.h, _p(c)))
        return c

    # -- per-stage
    def findFG(self, A, B, L, E, F, G):
        """findFG (MAIN:653-662), one call: returns the updated (F, G)"""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        F, G = f(F).copy(), f(G).copy()
        self._ck(self.lib.lqro_find_fg(self.h, _p(f(A)), _p(f(B)), _p(f(L)), _p(f(E)), _p(F), _p(G)))
        return F, G

    def createObstacleStep(self, points, F, G, P, x1, x2):
        """createObstacle (MAIN:683-688), one call: returns (q, Transform, Translate)"""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        points = f(points)
        q, T, tr = np.zeros_like(points), np.zeros((3, 3)), np.zeros(3)
        self._ck(self.lib.lqro_create_obstacle(self.h, len(points), _p(points), _p(f(F)), _p(f(G)), _p(f(P)), _p(f(x1)), _p(f(x2)), _p(T), _p(tr), _p(q)))
        return q, T, tr

    def createObstacle(self, xi, xj):
        q = np.zeros((self.H * self.P, 3))
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._ck(self.lib.lqro_candidates(self.h, _p(f(xi)), _p(f(xj)), _p(q)))
        return q

    def findReachableObstacle(self, q, xi, xj):
        q = np.ascontiguousarray(q, dtype=np.float64)
        idx = np.zeros(max(len(q), 1), dtype=np.int32)
        n = c_int()
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._ck(self.lib.lqro_reachable(self.h, _p(q), len(q), _p(f(xi)), _p(f(xj)), _p(idx), ctypes.byref(n)))
        return idx[: n.value].copy()

    def run_gjk(self, vrel, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        sq, it = c_double(), c_int()
        w, w1 = np.zeros(3), np.zeros(3)
        self._ck(self.lib.lqro_gjk(self.h, _p(np.ascontiguousarray(vrel, dtype=np.float64)), _p(pts), len(pts), ctypes.byref(sq),
                                   _p(w1), _p(w), ctypes.byref(it)))
        self.last_gjk_witness_vrel = w1
        return sq.value, w, it.value

    def convex_hull(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        cap = 2 * len(pts) + 64
        planes, verts = np.zeros((cap, 4)), np.zeros((cap, 3), dtype=np.int32)
        nf, deg = c_int(cap), c_int()
        self._ck(self.lib.lqro_convex_hull(self.h, _p(pts), len(pts), _p(planes), _p(verts), ctypes.byref(nf), ctypes.byref(deg)))
        return planes[: nf.value].copy(), verts[: nf.value].copy(), deg.value

    def convexHull(self, pts, vrel):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        d, best, nf = c_double(), c_int(), c_int()
        nrm = np.zeros(3)
        self._ck(self.lib.lqro_hull_escape(self.h, _p(pts), len(pts), _p(np.ascontiguousarray(vrel, dtype=np.float64)),
                                           ctypes.byref(d), _p(nrm), ctypes.byref(best), ctypes.byref(nf)))
        return d.value, nrm, best.value, nf.value

    def calculateNewV(self, planes, counts, goal):
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        batch, maxp = planes.shape[0], planes.shape[1]
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        out = np.zeros((batch, 3))
        self._ck(self.lib.lqro_new_velocity(self.h, batch, maxp, _p(planes), _p(counts), _p(goal), _p(out)))
        return out

    def agent_update(self, pos, vel, newVel, dt):
        pos = np.ascontiguousarray(pos, dtype=np.float32).copy()
        vel = np.ascontiguousarray(vel, dtype=np.float32).copy()
        newVel = np.ascontiguousarray(newVel, dtype=np.float32)
        self._ck(self.lib.lqro_agent_update(self.h, len(pos), _p(pos), _p(vel), _p(newVel), dt))
        return pos, vel

    def set_dynamics_class(self, cls, A, B, L, E):
        """dynamics class `cls` (0..15); class 0 == set_dynamics"""
        A, B, L, E = (np.ascontiguousarray(m, dtype=np.float64) for m in (A, B, L, E))
        self._ck(self.lib.lqro_set_dynamics_class(self.h, int(cls), _p(A), _p(B), _p(L), _p(E)))

    def set_agent_classes(self, cls):
        """class id of every agent (n_agents ints), or None for a single class"""
        if cls is None:
            self._ck(self.lib.lqro_set_agent_classes(self.h, None))
            return
        cls = np.ascontiguousarray(cls, dtype=np.int32)
        assert cls.shape == (self.N,)
        self._ck(self.lib.lqro_set_agent_classes(self.h, _p(cls)))

    def knn(self, x=None):
        """K nearest other agents of every owned agent on the device (AGENT:153-174 ordering); the table stays
        resident as the context's neighbour table.  x None = the state already on the device."""
        nbr = np.zeros((self.NL, self.K), dtype=np.int32)
        self._ck(self.lib.lqro_knn(self.h, _p(x), _p(nbr)))
        return nbr

    def knn_device(self, d_x=None, d_nbr=None, stream=None):
        self._ck(self.lib.lqro_knn_device(self.h, _p(d_x), _p(d_nbr), c_void_p(stream) if stream else None))

    def neighbors(self):
        nbr = np.zeros((self.NL, self.K), dtype=np.int32)
        self._ck(self.lib.lqro_get_neighbors(self.h, _p(nbr)))
        return nbr

    def round6(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.zeros_like(a)
        self._ck(self.lib.lqro_round6(self.h, _p(a), a.size, _p(out)))
        return out

    # -- multi-GPU
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = (ctypes.c_char * 128)()
        rc = lib.lqro_comm_unique_id(buf)
        if rc != 0:
            raise LqroError("lqro_comm_unique_id failed: %d" % rc)
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        buf = (ctypes.c_char * 128).from_buffer_copy(uid)
        self._ck(self.lib.lqro_comm_init(self.h, rank, world, buf))

    def comm_p2p_active(self):
        """True when the step's all-gather runs over NVLink peer memory (csrc/lqro_p2p.cuh), False when it is ncclAllGather"""
        return bool(self.lib.lqro_comm_p2p_active(self.h))

    def device_state(self):
        p = c_void_p()
        self._ck(self.lib.lqro_device_state(self.h, ctypes.byref(p)))
        return p.value

    def upload_local_state(self, x_local):
        self._ck(self.lib.lqro_upload_local_state(self.h, _p(x_local)))

    def allgather_state(self):
        self._ck(self.lib.lqro_allgather_state(self.h))

    def measure_peaks(self):
        a, b = c_double(), c_double()
        self._ck(self.lib.lqro_measure_fp64_peak(self.h, ctypes.byref(a)))
        self._ck(self.lib.lqro_measure_fp32_peak(self.h, ctypes.byref(b)))
        return a.value, b.value
