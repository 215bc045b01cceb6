"""Multi-GPU plumbing of the path: one process per GPU, torch.distributed for the exchange.

Two shardings exist and only one of them has a collective:

* sample sharding (BASELINE config C4): every rank scores the contiguous slice of the velocity-sample grid that
  ``b200lp_plan_shard(rank, count)`` selects and owns a local best ``(cost, global trajectory id)``. The global
  best is found with ONE small all-reduce: a zero-initialised int64 vector of 2*W entries in which rank r fills
  only slots [2r, 2r+1] with (bit pattern of its cost, id); the sum over ranks is an exact all-gather of 16 B per
  rank. Every rank then applies the reference's rule — minimum cost, ties to the LARGEST id
  (local_planner.cpp:460 keeps the last trajectory with `<=`) — so all ranks agree without a second exchange.
* fleet sharding (config C5): robots are independent, ranks own disjoint robot ranges, no collective.
"""
from __future__ import annotations

import struct

NONE_BITS = (1 << 63) - 1  # "no feasible trajectory": larger than the bits of any cost <= 9999999


def cost_to_bits(cost: float, best_id: int) -> int:
    """Non-negative doubles order like their bit patterns (sign bit clear), so int64 compares are exact."""
    if best_id < 0 or not (cost >= 0.0):
        return NONE_BITS
    return struct.unpack("<q", struct.pack("<d", float(cost)))[0]


def bits_to_cost(bits: int) -> float:
    return struct.unpack("<d", struct.pack("<q", int(bits)))[0]


def pick_best(pairs):
    """pairs: iterable of (cost_bits, id). Reference rule: min cost, ties -> largest id. -> (cost, id) or (-1.0, -1)."""
    best_bits, best_id = NONE_BITS, -1
    for bits, tid in pairs:
        if bits == NONE_BITS or tid < 0:
            continue
        if bits < best_bits or (bits == best_bits and tid > best_id):
            best_bits, best_id = int(bits), int(tid)
    if best_id < 0:
        return -1.0, -1
    return bits_to_cost(best_bits), best_id


def allreduce_best(local_cost: float, local_id: int, device=None, group=None):
    """The single collective of the sample-sharded path. Works on NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    vals = [0] * (2 * world)
    vals[2 * rank] = cost_to_bits(local_cost, local_id)
    vals[2 * rank + 1] = int(local_id)
    buf = torch.tensor(vals, dtype=torch.int64, device=device)  # one H2D copy
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    vals = buf.cpu().tolist()
    return pick_best((vals[2 * r], vals[2 * r + 1]) for r in range(world))


def shard_range(n: int, rank: int, count: int):
    """[lo, hi) of n units for shard `rank` of `count`: the equal-count split fleet sharding applies to robots. NOT the split
    of a sample-sharded cycle: b200lp_plan_shard cuts the sample grid at equal shares of the estimated pose count (and the
    exchange path moves the cuts with measured times); ask ``LocalPlanner.traj_count()`` for the trajectory-id range a
    shard really scored."""
    return n * rank // count, n * (rank + 1) // count


def attach_peer_exchange(planner, device=None, group=None, cloud_capacity: int = 0) -> bool:
    """Set the peer-memory exchanges up for ``planner`` (a LocalPlanner) in the current process group: gather every
    rank's 64-byte CUDA IPC handle with one all-gather and attach. ``cloud_capacity`` > 0 also reserves the row buffer of
    ``set_cloud_shared`` (points). Returns False (and leaves the planner on the all-reduce / per-rank upload path) when the
    box does not allow it; the decision is agreed on by all ranks. The attach ends with a collective, i.e. a barrier: no
    rank starts its first exchange before every rank is attached."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ok = 1
    try:
        if cloud_capacity:
            planner.peer_reserve_cloud(int(cloud_capacity))
        mine = planner.peer_export()
    except Exception:
        mine, ok = bytes(64), 0
    t = torch.tensor(list(mine) + [ok], dtype=torch.uint8, device=device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    rows = [bytes(g.cpu().tolist()) for g in gathered]
    if not all(r[64] for r in rows):
        return False
    try:
        planner.peer_attach(rank, [r[:64] for r in rows])
        ok = 1
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.cpu()[0]))


def resync_peer_exchange(planner, group=None) -> None:
    """After a failed plan_shard_exchange / set_cloud_shared on ANY rank: every rank calls this; sequence numbers, slots and
    shard cuts start over, fenced by two barriers."""
    import torch.distributed as dist

    dist.barrier(group)
    planner.peer_resync()
    dist.barrier(group)
