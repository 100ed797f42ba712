"""Host-side sharding of the path over ranks (one process per GPU).

Agents are partitioned in equal contiguous blocks, block r owned by rank r; each step every rank
contributes its block of the 16-double state estimate to one all-gather (no reduction, no second
exchange: a pair needs only the two agents' states, an agent's LP only its own planes).  On GPUs the
all-gather is ncclAllGather inside liblqro (lqro_allgather_state); this module holds the rank
arithmetic and a torch.distributed version of the same exchange, used by the CPU (gloo) tests.
"""
import numpy as np


def shard_bounds(n_agents, world, rank):
    """-> (local_begin, n_local).  Equal blocks are required by the in-place all-gather."""
    if n_agents % world != 0:
        raise ValueError("n_agents (%d) must be a multiple of the world size (%d)" % (n_agents, world))
    n_local = n_agents // world
    return rank * n_local, n_local


def local_slices(x, vpref, nbr, world, rank):
    """this rank's block of the per-agent inputs (neighbour ids stay global)"""
    lo, nl = shard_bounds(len(x), world, rank)
    return (np.ascontiguousarray(x[lo:lo + nl]), np.ascontiguousarray(vpref[lo:lo + nl]),
            np.ascontiguousarray(nbr[lo:lo + nl]))


def allgather_state(x_local, group=None):
    """torch.distributed all-gather of the per-rank state blocks -> (N, 16) numpy array on every rank."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(x_local))
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.numpy()


def gather_results(v_local, group=None):
    """all ranks' new velocities in agent order (used by tests / drivers that want the full vector)"""
    return allgather_state(v_local, group)


def local_normals(seed, n_agents, world, rank, per_agent=22):
    """Standard normals of one simulator-form update for this rank's agents.  The reference draws its noise from ONE sequential
    rand() stream over all quadrotors (MAIN:1408-1417), so every rank draws the whole step's per_agent*n_agents values
    (host work: microseconds per agent) and keeps its block; all ranks end on the same generator state.
    -> (normals (n_local, per_agent), new_seed)"""
    from .api import msvc_normals

    lo, nl = shard_bounds(n_agents, world, rank)
    allv, new_seed = msvc_normals(seed, per_agent * n_agents)
    return np.ascontiguousarray(allv.reshape(n_agents, per_agent)[lo:lo + nl]), new_seed
