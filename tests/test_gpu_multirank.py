"""N > 1 on real GPUs (needs >= 2 devices; skipped otherwise): two ranks, one GPU each, agents in two blocks,
ncclAllGather of the state inside lqro_step_sharded -- the assembled new velocities equal the one-GPU step bit for bit,
with host-provided neighbour lists and with the device neighbour search."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist

    from lqr_obstacles_b200 import Lqro, swarm
    from lqr_obstacles_b200 import dist as ldist
    from lqr_obstacles_b200.api import FLAG_DEVICE_KNN

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    g = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))
    x, vpref = swarm.lattice_swarm((8, 8, 4))
    x[:, :3] *= 0.6
    K = 16
    nbr = swarm.knn(x[:, :3], K)
    xl, vl, nl = ldist.local_slices(x, vpref, nbr, world, rank)
    lo, n_local = ldist.shard_bounds(len(x), world, rank)
    out = {}
    for name, flags, nb in (("host", 0, nl), ("knn", FLAG_DEVICE_KNN, None)):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")  # one NCCL unique id per communicator
        if rank == 0:
            uid = torch.frombuffer(bytearray(Lqro.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx = Lqro(device=rank, n_agents=len(x), local_begin=lo, n_local=n_local, max_neighbors=K, horizon=20, flags=flags)
        ctx.set_dynamics(g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"])
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        v = ctx.step_sharded(np.ascontiguousarray(xl), np.ascontiguousarray(vl), None if nb is None else np.ascontiguousarray(nb))
        t = torch.from_numpy(v).cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        out[name] = torch.cat(parts).cpu().numpy()
        ctx.close()
    if rank == 0:
        one = Lqro(device=0, n_agents=len(x), max_neighbors=K, horizon=20)
        one.set_dynamics(g["dyn_A"], g["dyn_B"], g["dyn_L"], g["dyn_E"])
        v_ref = one.step(x, vpref, nbr)
        one.close()
        np.save(os.path.join(outdir, "ok.npy"), np.array([np.array_equal(out["host"], v_ref), np.array_equal(out["knn"], v_ref)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_match_one(tmp_path, product_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.load(os.path.join(str(tmp_path), "ok.npy")).all()
