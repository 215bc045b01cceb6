"""world_size-2 (and 4) gloo runs of the sample-shard argmin exchange (dddmr_navigation_b200/dist.py), on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dddmr_navigation_b200 import dist as lpdist
from tests.helpers import reference_argmin


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, costs, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = lpdist.shard_range(len(costs), rank, world)
    local = costs[lo:hi]
    lid = reference_argmin(local)  # what b200lp_plan_shard returns for the slice, as a GLOBAL id
    lcost = float(local[lid]) if lid >= 0 else -1.0
    gid = lo + lid if lid >= 0 else -1
    cost, best = lpdist.allreduce_best(lcost, gid)
    out_q.put((rank, cost, best))
    dist.destroy_process_group()


def _run(world, costs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, costs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_argmin_matches_reference_rule(world):
    rng = np.random.default_rng(world)
    costs = rng.uniform(0.5, 5.0, 1001)
    costs[rng.integers(0, 1001, 300)] = -1.0          # rejected trajectories
    costs[[17, 400, 999]] = 0.25                      # a three-way tie spanning shards: the LAST id must win
    want = reference_argmin(costs)
    assert want == 999
    for rank, cost, best in _run(world, costs):
        assert best == want and cost == costs[want], (rank, cost, best)


def test_sharded_argmin_all_rejected_and_empty_shard():
    costs = np.full(3, -1.0)  # 3 trajectories over 2 ranks, none feasible
    for rank, cost, best in _run(2, costs):
        assert best == -1 and cost == -1.0
    costs = np.array([2.0])   # rank 0's slice is empty
    for rank, cost, best in _run(2, costs):
        assert best == 0 and cost == 2.0


def test_cost_bits_order_like_costs():
    xs = np.sort(np.concatenate([[0.0, 5e-324, 9999999.0], np.random.default_rng(0).uniform(0, 1e7, 1000)]))
    bits = [lpdist.cost_to_bits(float(x), 1) for x in xs]
    assert bits == sorted(bits) and all(b < lpdist.NONE_BITS for b in bits)
    assert lpdist.cost_to_bits(-1.0, -1) == lpdist.NONE_BITS and lpdist.cost_to_bits(float("nan"), 3) == lpdist.NONE_BITS
    assert lpdist.pick_best([(lpdist.cost_to_bits(1.0, 4), 4), (lpdist.cost_to_bits(1.0, 9), 9), (lpdist.NONE_BITS, -1)]) == (1.0, 9)
