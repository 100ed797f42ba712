"""world_size-2 test of the N>1 host path on CPU (gloo): rank arithmetic, the state all-gather and the
assembly of per-rank results.  The per-pair compute is done by the CPU oracle here (no GPU in this
container); on GPUs the same sharding drives liblqro with ncclAllGather (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from _libs import Oracle
    from lqr_obstacles_b200 import dist as ldist
    from lqr_obstacles_b200 import swarm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
    x, vpref = swarm.lattice_swarm((4, 4, 2))
    x[:, :3] *= 0.6
    nbr = swarm.knn(x[:, :3], 6)
    xl, vl, nl = ldist.local_slices(x, vpref, nbr, world, rank)
    lo, n_local = ldist.shard_bounds(len(x), world, rank)
    x_all = ldist.allgather_state(xl)            # the per-step exchange
    assert np.array_equal(x_all, x)
    O = Oracle()
    NF, T, _ = O.horizon(g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"], 20)
    sph = O.spheres()
    full = np.full((len(x), nbr.shape[1]), -1, dtype=np.int32)   # only the owned agents have neighbours here
    full[lo:lo + n_local] = nl
    v = O.step(NF, T, sph, x_all, vpref, full)[0][lo:lo + n_local]
    v_all = ldist.gather_results(np.ascontiguousarray(v))
    if rank == 0:
        v_ref = O.step(NF, T, sph, x, vpref, nbr)[0]
        np.save(os.path.join(outdir, "ok.npy"), np.array([np.array_equal(v_all, v_ref)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(os.path.join(str(tmp_path), "ok.npy"))[0]


def _sim_worker(rank, world, port, outdir):
    """simulator-form update sharded over two ranks (the CPU build of the update kernels' source does the per-agent work
    here): each rank updates its own quads with its block of the step's normals; the assembled result equals the
    one-process update bit for bit and both ranks end on the same rand() state"""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from _libs import SimHost
    from lqr_obstacles_b200 import dist as ldist
    from lqr_obstacles_b200.api import msvc_normals

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(HERE, "golden", "sim_update.npz"))
    sh = SimHost()
    n, t = 6, 90
    lo, nl = ldist.shard_bounds(n, world, rank)
    seed0 = int(g["seeds"][t])
    nrm, seed1 = ldist.local_normals(seed0, n, world, rank)
    assert np.array_equal(nrm, g["s%d_normals" % t][lo:lo + nl])
    st = {k: np.ascontiguousarray(g["s%d_before_%s" % (t, k)][lo:lo + nl]) for k in ("x", "Rot", "P", "xTrue", "RotTrue", "pGoal")}
    o = sh.update(g["model"], st, np.ascontiguousarray(g["s%d_newV" % t][lo:lo + nl]), nrm)
    x_all = ldist.allgather_state(o["x"])          # the exchange the next avoidance step starts with
    xt_all = ldist.gather_results(o["xTrue"])
    ok = [np.array_equal(x_all, g["s%d_after_x" % t]), np.array_equal(xt_all, g["s%d_after_xTrue" % t]),
          seed1 == msvc_normals(seed0, 22 * n)[1]]
    if t + 1 < len(g["seeds"]):
        ok.append(seed1 == int(g["seeds"][t + 1]))
    np.save(os.path.join(outdir, "ok%d.npy" % rank), np.array(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_simulator_update_matches_the_reference(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_sim_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(os.path.join(str(tmp_path), "ok%d.npy" % r)).all()


def test_shard_bounds():
    sys.path.insert(0, os.path.dirname(HERE))
    from lqr_obstacles_b200.dist import shard_bounds

    assert [shard_bounds(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 8192)]
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
