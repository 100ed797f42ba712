"""End-to-end time of the avoidance step from pinned HOST buffers (lqro_step_sharded on one GPU: state, preferred velocities and
neighbour table up, new velocities down, every step), with the neighbour table uploaded next to the pair kernel (default) and in
front of it (LQRO_NBR_STREAM_MIN_PAIRS=0), and the device-resident step beside them.   usage: python tools/quick_e2e.py [dims K steps]"""
import sys, os, time, numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
from lqr_obstacles_b200 import Lqro, swarm
g = np.load(os.path.join(root, 'tests', 'golden', 'ref_vectors.npz'))
dims = tuple(int(v) for v in sys.argv[1].split('x')) if len(sys.argv) > 1 else (64, 32, 32)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
x, vp = swarm.lattice_swarm(dims); nb = swarm.knn(x[:, :3], K)
N = len(x)
hx, hv, hn = torch.from_numpy(x).pin_memory(), torch.from_numpy(vp).pin_memory(), torch.from_numpy(nb).pin_memory()
out = {}
for mode in ("streamed", "in_front", "streamed", "in_front"):
    os.environ["LQRO_NBR_STREAM_MIN_PAIRS"] = "0" if mode == "in_front" else str(1 << 17)
    ctx = Lqro(n_agents=N, max_neighbors=K, horizon=20); ctx.set_dynamics(g['dyn_A'], g['dyn_B'], g['dyn_L'], g['dyn_E'])
    ho = torch.zeros(N, 3, dtype=torch.float64).pin_memory()
    for _ in range(3): ctx.step_sharded(hx.data_ptr(), hv.data_ptr(), hn.data_ptr(), ho.data_ptr())
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ctx.step_sharded(hx.data_ptr(), hv.data_ptr(), hn.data_ptr(), ho.data_ptr())
        ts.append((time.perf_counter() - t0) * 1e3)
    dev = ctx.last_step_ms()
    print(mode, "streamed" if ctx.last_step_nbr_streamed() else "in front", "e2e ms/step median %.4f min %.4f mean %.4f" % (np.median(ts), min(ts), np.mean(ts)),
          "kernels of the last step", {k: round(v, 4) for k, v in dev.items()}, ctx.last_step_detail_ms(), ctx.counters(), flush=True)
    out.setdefault(mode, ho.numpy().copy())
    st = ctx.pair_results()["status"]
    out.setdefault(mode + "_st", st.copy())
    ctx.close()
print("identical new velocities:", bool(np.array_equal(out["streamed"], out["in_front"])), "identical statuses:", bool(np.array_equal(out["streamed_st"], out["in_front_st"])))
