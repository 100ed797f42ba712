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


def test_shard_bounds():
    sys.path.insert(0, os.path.dirname(HERE))
    from lqr_obstacles_b200.dist import shard_bounds

    assert [shard_bounds(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 8192)]
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
